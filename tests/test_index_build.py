"""The fixture-side index builder writes byte-identical files to the reference's `FASTQuick index`
(BwtIndexer::BuildIndex): .pac/.rpac/.bwt/.rbwt/.sa/.rsa/.ann/.amb/.gc/.rollhash and the flank FASTA."""
import hashlib
import os
import subprocess

import pytest

import fx


def _md5(path):
    h = hashlib.md5()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 24), b""):
            h.update(chunk)
    return h.hexdigest()


def test_builder_matches_reference_index(ref_required):
    if not os.path.exists(fx.REF_BIN):
        pytest.skip("FASTQuick_ref not built")
    idx = fx.SynthIndex("ibuild", n_long=12, n_short=60, n_x=3, n_y=2, seed=77, with_rollhash=True)
    env = dict(os.environ, PATH=os.path.join(fx.REPO, "oracle", "_ref") + ":" + os.environ["PATH"])
    ref_prefix = os.path.join(idx.dir, "ref")
    if not os.path.exists(ref_prefix + ".FASTQuick.fa.rsa"):
        subprocess.check_call([fx.REF_BIN, "index", "--predefinedVCF", "markers.vcf", "--dbsnpVCF", "dbsnp.vcf", "--ref", "genome.fa",
                               "--out_prefix", "ref", "--var_long", "12", "--var_short", "60"], cwd=idx.dir, env=env,
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    for ext in ("", ".pac", ".rpac", ".bwt", ".rbwt", ".sa", ".rsa", ".ann", ".amb", ".gc", ".bed", ".SelectedSite.vcf", ".rollhash"):
        assert _md5(idx.prefix + ext) == _md5(ref_prefix + ".FASTQuick.fa" + ext), ext


def test_builder_matches_reference_index_with_ambiguous_bases(ref_required):
    """Flanks holding N, an IUPAC code and lower-case bases: the reference's `FASTQuick index` run on such a genome, and the
    fixture builder fed the flank FASTA that run extracted, leave identical files (pins the nt4 code 4 'bleed' of
    AddSeq2HashCore's rolling k-mer and Fa2Pac's .amb holes + lrand48 substitutes)."""
    if not os.path.exists(fx.REF_BIN):
        pytest.skip("FASTQuick_ref not built")
    import ctypes as C
    import shutil
    src = fx.SynthIndex("ibuild", n_long=12, n_short=60, n_x=3, n_y=2, seed=77, with_rollhash=True)
    d = os.path.join(fx.CACHE, "ibuild_n")
    os.makedirs(d, exist_ok=True)
    ref_prefix = os.path.join(d, "ref")
    if not os.path.exists(os.path.join(d, ".done")):
        for f in ("genome.fa", "genome.fa.fai", "markers.vcf", "dbsnp.vcf"):
            shutil.copy(os.path.join(src.dir, f), os.path.join(d, f))
        g = bytearray(open(os.path.join(d, "genome.fa"), "rb").read())
        body = g.index(b"\n") + 1                                   # chromosome "1": 60 bases per line
        at = lambda pos1: body + (pos1 - 1) + (pos1 - 1) // 60
        for pos1, ch in ((1990, b"N"), (2050, b"N"), (2051, b"N"), (2052, b"N"), (2100, b"R"), (5010, b"n"), (4990, b"a"), (8020, b"N")):
            if ch == b"a": ch = bytes([g[at(pos1)]]).lower()
            g[at(pos1)] = ch[0]
        open(os.path.join(d, "genome.fa"), "wb").write(bytes(g))
        env = dict(os.environ, PATH=os.path.join(fx.REPO, "oracle", "_ref") + ":" + os.environ["PATH"])
        subprocess.check_call([fx.REF_BIN, "index", "--predefinedVCF", "markers.vcf", "--dbsnpVCF", "dbsnp.vcf", "--ref", "genome.fa",
                               "--out_prefix", "ref", "--var_long", "12", "--var_short", "60"], cwd=d, env=env,
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        lib = fx.host_lib()
        rc = lib.fqb_index_from_flank_fasta((ref_prefix + ".FASTQuick.fa").encode(), os.path.join(d, "mine.FASTQuick.fa").encode(), 1)
        assert rc == 0, lib.fqb_last_error()
        open(os.path.join(d, ".done"), "w").close()
    flank_text = open(ref_prefix + ".FASTQuick.fa").read()
    assert any(c in flank_text for c in "Nn"), "the ambiguous bases did not reach a flank"
    assert int(open(ref_prefix + ".FASTQuick.fa.amb").readline().split()[2]) > 0
    for ext in (".pac", ".rpac", ".bwt", ".rbwt", ".sa", ".rsa", ".ann", ".amb", ".rollhash"):
        assert _md5(ref_prefix + ".FASTQuick.fa" + ext) == _md5(os.path.join(d, "mine.FASTQuick.fa") + ext), ext
