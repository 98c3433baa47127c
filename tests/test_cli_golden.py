"""End-to-end record identity with the reference CLI.

tests/golden/cli/ holds the output files of the UNMODIFIED reference CLI on its
own FASTA fixtures for all five formats (oracle/gen_cli_golden.py; the default
run on test.fa equals the reference's own repeat.tab).

  * CPU test: the host glue alone -- FASTA loading, post-processing chain and
    writers -- fed with raw records from the CPU oracle's strict scan.
  * GPU test: the real thing, `bwt.main([...])`, byte for byte.
"""
import os

import numpy as np
import pytest

CLI = os.path.join(os.path.dirname(__file__), "golden", "cli")
CASES = [
    ("test.fa", [], "test"),
    ("test.fa", ["--tier1"], "test_tier1"),
    ("test2.fa", [], "test2"),
    ("test2.fa", ["--max-motif-len", "12"], "test2_max-motif-len_12"),
    ("test_long_motif.fa", [], "test_long_motif"),
    ("test_synthetic.fasta", ["--flank-trim", "0"], "test_synthetic_flank-trim_0"),
]
# a FASTA of our own through the reference CLI (oracle/gen_cli_golden.py: write_syn_mixed): insertions / deletions in
# closely spaced arrays, lower-case stretches, an N block, IUPAC codes, a contig too short to trim, ragged lines.
CPU_CASES = CASES + [("syn_mixed.fa", [], "syn_mixed")]       # the host-glue test and the GPU CLI test run all of these
FORMATS = ["strfinder", "bed", "vcf", "trf_table", "trf_dat"]


def _expected(tag, fmt):
    with open(os.path.join(CLI, f"{tag}.{fmt}.txt")) as f:
        return f.read()


@pytest.mark.parametrize("fa,flags,tag", CPU_CASES, ids=[c[2] for c in CPU_CASES])
def test_host_glue_reproduces_reference_outputs(oracle, tmp_path, fa, flags, tag):
    import bwt_algorithm_b200  # noqa: F401
    from bwt_algorithm_b200 import finders
    from bwt_algorithm_b200.pipeline import TandemRepeatFinder

    trim = 0 if "--flank-trim" in flags else 30
    finder = TandemRepeatFinder(os.path.join(CLI, fa), flank_trim=trim,
                                max_motif_length=12 if "--max-motif-len" in flags else 9)
    seqs = finder.load_reference()
    raw = []
    if "--tier1" not in flags:
        for chrom, seq in seqs.items():
            text = (seq + "$").encode()
            eff = max(120, min(len(seq) // 3, 1000))
            rows = oracle.strict_scan(text, 1, eff, 0, 3)
            raw.extend(finders.strict_records(np.frombuffer(text, np.uint8), rows, chrom, 0))
    final, _dups = finder._postprocess(raw, lambda m: None)
    # the row-oriented chain per contig (what the CLI runs) yields the same calls
    import dataclasses

    from bwt_algorithm_b200 import pipeline, rowchain

    by_rows = []
    if "--tier1" not in flags and pipeline._distinct_sort_names(seqs):
        cfg = finder._config(True, True)
        for chrom, seq, c, left, right in finder._finish_tasks([(ch, sq, cfg) for ch, sq in seqs.items()]):
            text = (seq + "$").encode()
            rows = oracle.strict_scan(text, 1, max(120, min(len(seq) // 3, 1000)), 0, 3)
            by_rows.extend(rowchain.finish_rows(pipeline._contig_finder(chrom, seq, c, left, right), chrom,
                                                np.frombuffer(text, np.uint8), rows)[0])
        by_rows.sort(key=finder._repeat_sort_key)
        assert [dataclasses.astuple(r) for r in by_rows] == [dataclasses.astuple(r) for r in final]
    for fmt in FORMATS:
        out = tmp_path / f"{tag}.{fmt}"
        finder.save_results(final, str(out), fmt)
        assert out.read_text() == _expected(tag, fmt), f"{tag} ({fmt}) differs from the reference CLI output"


@pytest.mark.gpu
@pytest.mark.parametrize("jobs", ["-1", "4"])
@pytest.mark.parametrize("fa,flags,tag", CPU_CASES, ids=[c[2] for c in CPU_CASES])
def test_cli_byte_identical_on_gpu(tmp_path, capsys, fa, flags, tag, jobs):
    import bwt_algorithm_b200  # noqa: F401
    from bwt_algorithm_b200 import bwt

    for fmt in FORMATS:
        out = tmp_path / f"{tag}.{fmt}"
        bwt.main([os.path.join(CLI, fa), "-o", str(out), "--format", fmt, "--jobs", jobs] + flags)
        assert out.read_text() == _expected(tag, fmt), f"{tag} ({fmt}, jobs {jobs}) differs"
    capsys.readouterr()


@pytest.mark.gpu
def test_cli_chain_pool_on_gpu(tmp_path, capsys, monkeypatch):
    """Several contigs on one GPU with --jobs 4: detection in this process (CUDA context up), every contig's
    chain in a spawned CPU worker (pipeline._run_with_chain_pool, forced on for a small FASTA).  Byte-identical
    output; the pool path really ran."""
    import bwt_algorithm_b200  # noqa: F401
    from bwt_algorithm_b200 import bwt, pipeline, sharding

    fa, flags, tag = next(c for c in CASES if c[0] == "test2.fa" and "--tier1" not in c[1])
    monkeypatch.setattr(pipeline, "CHAIN_POOL_MIN_BASES", 0)
    monkeypatch.setattr(sharding, "worker_processes", lambda n_tasks, n_jobs: 1)
    used = []
    real = pipeline._run_with_chain_pool
    monkeypatch.setattr(pipeline, "_run_with_chain_pool", lambda tasks, procs: used.append(procs) or real(tasks, procs))
    for fmt in ("bed", "strfinder"):
        out = tmp_path / f"{tag}.{fmt}"
        bwt.main([os.path.join(CLI, fa), "-o", str(out), "--format", fmt, "--jobs", "4"] + flags)
        assert out.read_text() == _expected(tag, fmt), f"{tag} ({fmt}, chain pool) differs"
    assert used and all(p >= 2 for p in used)
    capsys.readouterr()


@pytest.mark.gpu
def test_reference_unit_tests_on_gpu():
    """The reference's own assertions (tests/test_repeat_outputs.py:36-62) through the drop-in."""
    import bwt_algorithm_b200  # noqa: F401
    from bwt_algorithm_b200.bwt import TandemRepeatFinder

    finder = TandemRepeatFinder(os.path.join(CLI, "test2.fa"), show_progress=False, max_motif_length=12)
    finder.load_reference()
    repeats = finder.find_tandem_repeats(enable_tier1=True, enable_tier2=True, enable_tier3=False)
    by = {}
    for r in repeats:
        by.setdefault(r.chrom, []).append(r)
    r1 = by["test1_PERFECT_7mer_5copies"][0]
    assert (r1.start, r1.end, r1.motif, r1.copies, r1.variations) == (30, 65, "TCATCGG", 5.0, None)
    assert len(by["test4_INTERRUPTED_7mer_11copies"]) == 1
    r4 = by["test4_INTERRUPTED_7mer_11copies"][0]
    assert r4.variations is not None and set(r4.variations) == {"6:5:C>A", "10:6:G>A", "11:0:ins(G)"}
    motifs = {r.motif for r in by["test6_NESTED_long20_short4"]}
    assert "TGCTGATCGTAGCTAGCTGA" in motifs and "TGCT" in motifs and "CTGA" not in motifs
    assert len(by["test12_LONG_IMPERFECT_indel"]) == 1
    r12 = by["test12_LONG_IMPERFECT_indel"][0]
    assert r12.variations and any(v.startswith("9:10:del(") for v in r12.variations)


# ---- BASELINE configs[1]: the 150 kb synthetic contig `synC` ----------------------
SYNC_FASTA_MD5 = "f1e57a0252738b17088fa3ca6dddedf9"
SYNC_BED_MD5 = "d984e3dc49b02c179ee785ac12bd9330"      # the reference CLI, --jobs 0 --format bed (565 lines)


def _write_sync(path):
    import hashlib

    from tests.util import gen_contig

    seq = gen_contig(150_000, 42).tobytes().decode()
    with open(path, "w") as fh:
        fh.write(">synC\n")
        for i in range(0, len(seq), 80):
            fh.write(seq[i:i + 80] + "\n")
    with open(path, "rb") as fh:
        assert hashlib.md5(fh.read()).hexdigest() == SYNC_FASTA_MD5, "generator drifted from SURVEY Appendix B"


@pytest.mark.gpu
@pytest.mark.parametrize("jobs", ["0", "-1"])
def test_cli_synC_bed_md5_on_gpu(tmp_path, capsys, jobs):
    """`python bwt.py synC.fa --jobs 0 --format bed` == the reference CLI's output, by md5."""
    import hashlib

    import bwt_algorithm_b200  # noqa: F401
    from bwt_algorithm_b200 import bwt

    fa, out = tmp_path / "synC.fa", tmp_path / "synC.bed"
    _write_sync(fa)
    bwt.main([str(fa), "-o", str(out), "--jobs", jobs, "--format", "bed"])
    capsys.readouterr()
    data = out.read_bytes()
    assert data.count(b"\n") == 565
    assert hashlib.md5(data).hexdigest() == SYNC_BED_MD5


@pytest.mark.gpu
def test_cli_segment_pool_on_gpu(tmp_path, capsys, monkeypatch):
    """One contig, --jobs 4: detection in this process, the contig's chain cut into pieces for spawned CPU workers
    (pipeline._run_with_segment_pool, forced on for the 150 kb contig).  The reference CLI's bytes, by md5."""
    import hashlib

    import bwt_algorithm_b200  # noqa: F401
    from bwt_algorithm_b200 import bwt, pipeline, rowchain, sharding

    fa, out = tmp_path / "synC.fa", tmp_path / "synC.bed"
    _write_sync(fa)
    monkeypatch.setattr(pipeline, "SEGMENT_POOL_MIN_BASES", 0)
    monkeypatch.setattr(rowchain, "SEGMENT_MIN_ROWS", 500)
    monkeypatch.setattr(sharding, "worker_processes", lambda n_tasks, n_jobs: 1)
    used = []
    real = pipeline._finish_rows_in_pool
    monkeypatch.setattr(pipeline, "_finish_rows_in_pool", lambda *a, **k: used.append(1) or real(*a, **k))
    bwt.main([str(fa), "-o", str(out), "--jobs", "4", "--format", "bed"])
    capsys.readouterr()
    assert used
    assert hashlib.md5(out.read_bytes()).hexdigest() == SYNC_BED_MD5


@pytest.mark.gpu
def test_cli_fresh_process_answers_without_importing_torch(tmp_path):
    """`python bwt.py ...` in a fresh interpreter: the detector call runs on the library's own buffers (lean.py) and
    torch is never imported -- same bytes as the reference CLI for synC (md5) and for a multi-contig fixture in two
    formats, with --jobs 4 (chain pool) and --jobs -1; with BWTK_LEAN=0 the same process goes through torch."""
    import hashlib
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    fa = tmp_path / "synC.fa"
    _write_sync(fa)
    prog = ("import sys; sys.path.insert(0, {root!r}); import bwt; bwt.main({argv!r}); "
            "print('TORCH_IMPORTED', 'torch' in sys.modules)")

    def run(argv, **env):
        res = subprocess.run([sys.executable, "-c", prog.format(root=root, argv=argv)], capture_output=True, text=True,
                             env=dict(os.environ, **env), timeout=300)
        assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
        assert "ERROR" not in res.stdout, res.stdout[-2000:]
        return res.stdout

    for jobs in ("0", "-1"):
        out = tmp_path / f"synC_{jobs}.bed"
        stdout = run([str(fa), "-o", str(out), "--jobs", jobs, "--format", "bed"])
        assert "TORCH_IMPORTED False" in stdout
        assert hashlib.md5(out.read_bytes()).hexdigest() == SYNC_BED_MD5
    fixture, flags, tag = next(c for c in CASES if c[0] == "test2.fa" and not c[1])
    for fmt in ("bed", "strfinder"):
        out = tmp_path / f"{tag}.{fmt}"
        stdout = run([os.path.join(CLI, fixture), "-o", str(out), "--format", fmt, "--jobs", "4"] + flags)
        assert "TORCH_IMPORTED False" in stdout
        assert out.read_text() == _expected(tag, fmt)
    out = tmp_path / "synC_torch.bed"
    stdout = run([str(fa), "-o", str(out), "--jobs", "0", "--format", "bed"], BWTK_LEAN="0")
    assert "TORCH_IMPORTED True" in stdout
    assert hashlib.md5(out.read_bytes()).hexdigest() == SYNC_BED_MD5
