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
# a FASTA of our own (written by write_syn_mixed below, deterministic): what the reference's fixtures do not hold
GENERATED = [("syn_mixed.fa", [])]


def write_syn_mixed(path):
    """Four contigs: closely spaced arrays with substitutions, insertions and deletions (lower-case stretches);
    an N block and IUPAC codes; a contig too short for flank trimming; an all-lower-case contig.  Ragged line
    widths, blank lines, trailing blanks, a header with a description."""
    import numpy as np

    sys.path.insert(0, os.path.dirname(HERE))
    from tests.util import gen_contig, gen_contig_indel

    def lines(seq, width):
        return "\n".join(seq[i:i + width] for i in range(0, len(seq), width))

    a = gen_contig_indel(9000, 101).tobytes().decode()
    a = "".join(a[i:i + 500].lower() if (i // 500) % 2 else a[i:i + 500] for i in range(0, len(a), 500))
    b = gen_contig(7000, 102)
    b[2000:2300] = ord("N")
    rng = np.random.default_rng(5)
    b[rng.integers(0, 7000, 10)] = np.frombuffer(b"RYKMSW", np.uint8)[rng.integers(0, 6, 10)]
    b = b.tobytes().decode()
    c = "ACGT" * 12 + "AC"
    d = gen_contig_indel(4000, 103).tobytes().decode().lower()
    with open(path, "w") as fh:
        fh.write(">mix1 indel-rich arrays, lower-case stretches\n" + lines(a, 70) + "\n\n")
        fh.write(">mix2_N  N block and IUPAC codes \n" + lines(b, 61) + "  \n")
        fh.write(">mix3_short\n" + c + "\n")
        fh.write(">mix4_lower\n" + lines(d, 80) + "\n")


def main():
    os.makedirs(OUT, exist_ok=True)
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1", NUMBA_CACHE_DIR="/tmp/numba_cache_golden")
    only = sys.argv[1] if len(sys.argv) > 1 else None
    with tempfile.TemporaryDirectory() as tmp:
        for fa, flags in CASES + GENERATED:
            if only and fa != only:
                continue
            if (fa, flags) in GENERATED:
                write_syn_mixed(os.path.join(OUT, fa))
            else:
                shutil.copy(os.path.join(REF_DIR, fa), os.path.join(OUT, fa))
            for fmt in FORMATS:
                tag = fa.split(".")[0] + ("_" + "_".join(f.strip("-") for f in flags) if flags else "")
                out = os.path.join(OUT, f"{tag}.{fmt}.txt")
                for jobs in ("-1",):
                    cmd = [sys.executable, os.path.join(REF_DIR, "bwt.py"), os.path.join(OUT, fa), "-o", out,
                           "--format", fmt, "--jobs", jobs] + flags
                    subprocess.run(cmd, check=True, cwd=tmp, env=env, stdout=subprocess.DEVNULL)
                print(out, os.path.getsize(out))
    if only and only != "test.fa":
        return
    ref_tab = open(os.path.join(REF_DIR, "repeat.tab")).read()
    assert ref_tab == open(os.path.join(OUT, "test.strfinder.txt")).read(), "repeat.tab differs from a fresh run"
    print("repeat.tab == fresh default run on test.fa")


if __name__ == "__main__":
    main()
