#!/usr/bin/env python3
"""Generate tests/golden/cli/: outputs of the UNMODIFIED reference CLI
(`python /root/reference/bwt.py`) on its own FASTA fixtures, for every output
format.  Build-container only (needs /root/reference)."""
import os
import shutil
import subprocess
import sys
import tempfile

REF_DIR = os.environ.get("BWT_REFERENCE_DIR", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "cli")

CASES = [
    ("test.fa", []),                       # default flags: must equal the reference's own repeat.tab
    ("test.fa", ["--tier1"]),              # header-only output
    ("test2.fa", []),
    ("test2.fa", ["--max-motif-len", "12"]),
    ("test_long_motif.fa", []),
    ("test_synthetic.fasta", ["--flank-trim", "0"]),
]
FORMATS = ["strfinder", "bed", "vcf", "trf_table", "trf_dat"]


def main():
    os.makedirs(OUT, exist_ok=True)
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1", NUMBA_CACHE_DIR="/tmp/numba_cache_golden")
    with tempfile.TemporaryDirectory() as tmp:
        for fa, flags in CASES:
            shutil.copy(os.path.join(REF_DIR, fa), os.path.join(OUT, fa))
            for fmt in FORMATS:
                tag = fa.split(".")[0] + ("_" + "_".join(f.strip("-") for f in flags) if flags else "")
                out = os.path.join(OUT, f"{tag}.{fmt}.txt")
                for jobs in ("-1",):
                    cmd = [sys.executable, os.path.join(REF_DIR, "bwt.py"), os.path.join(OUT, fa), "-o", out,
                           "--format", fmt, "--jobs", jobs] + flags
                    subprocess.run(cmd, check=True, cwd=tmp, env=env, stdout=subprocess.DEVNULL)
                print(out, os.path.getsize(out))
    ref_tab = open(os.path.join(REF_DIR, "repeat.tab")).read()
    assert ref_tab == open(os.path.join(OUT, "test.strfinder.txt")).read(), "repeat.tab differs from a fresh run"
    print("repeat.tab == fresh default run on test.fa")


if __name__ == "__main__":
    main()
