"""Fixture of BASELINE.json configs[0] -- the reference's own example (example/example.sh, ctest Test2 at CMakeLists.txt:43-55):
ERR013170 1000-line FASTQ pair vs ref.test.fa with the hapmap.test.vcf.gz markers.

Run in the build container (needs /root/reference and oracle/_ref):   python tests/golden/make_example.py
It runs the REFERENCE's `index` and `align` on the example inputs and stores, under tests/golden/example/:
  * the inputs the align stage reads: the two FASTQ files, fq.test.list, the index files `FASTQuick_ref index` wrote
    (all but the 3 GiB .rollhash, which the product rebuilds on the device from the flank text) and ref.test.fa.{fai,amb};
  * ref_out.*: every file `FASTQuick_ref align --fq_list fq.test.list` wrote (12 statistics files + the BAM).
tests/test_gpu_example.py runs `FASTQuick_b200 align` on the same inputs and compares file by file / record by record."""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/example"
REF_BIN = os.path.join(REPO, "oracle", "_ref", "FASTQuick_ref")
OUT = os.path.join(HERE, "example")


def main():
    work = tempfile.mkdtemp(prefix="fqb_example_")
    for f in os.listdir(REF):
        if not f.endswith(".sh"):
            shutil.copy(os.path.join(REF, f), work)
    env = dict(os.environ, PATH=os.path.join(REPO, "oracle", "_ref") + os.pathsep + os.environ["PATH"])   # the bcftools stand-in
    subprocess.check_call([REF_BIN, "index", "--siteVCF", "hapmap.test.vcf.gz", "--dbsnpVCF", "dbsnp.test.vcf.gz", "--ref", "ref.test.fa",
                           "--out_prefix", "test_out_ref"], cwd=work, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    subprocess.check_call([REF_BIN, "align", "--fq_list", "fq.test.list", "--index_prefix", "test_out_ref", "--out_prefix", "ref_out"],
                          cwd=work, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    os.makedirs(OUT, exist_ok=True)
    keep = [f for f in os.listdir(work) if (f.startswith("test_out_ref.") and not f.endswith(".rollhash")) or f.startswith("ref_out.")
            or f.endswith(".fastq.gz") or f in ("fq.test.list", "ref.test.fa.fai", "ref.test.fa.amb")]
    for f in sorted(keep):
        shutil.copy(os.path.join(work, f), OUT)
    print("wrote %d files to %s" % (len(keep), OUT))
    shutil.rmtree(work)


if __name__ == "__main__":
    sys.exit(main())
