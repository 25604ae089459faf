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
