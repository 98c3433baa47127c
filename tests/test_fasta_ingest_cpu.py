"""FASTA ingest: the native two-pass parser (csrc/fasta.cu, bwtk_fasta_index /
bwtk_fasta_sequence) against the reference's line loop (bwt.py:3713-3756) -- byte-identical `sequences`,
`full_sequences` and `trim_offsets` on the reference's fixture FASTAs and on adversarial files."""
import glob
import os

import pytest

CLI = os.path.join(os.path.dirname(__file__), "golden", "cli")

ADVERSARIAL = {
    "lower_and_iupac": b">c1 some description\nacgtnRYKM\nacgtACGTnn\n>c2\nNNNNacgu\n",
    "crlf": b">c1\r\nACGT\r\nacgt\r\n>c2 x\r\nGGCC\r\n",
    "lone_cr": b">c1\rACGT\racgt\r>c2\rTTTT",
    "no_final_newline": b">only\nACGTACGTAC",
    "junk_before_first_header": b"ACGT\n\n  \nGG\n>c1\nAC\nGT\n",
    "interior_whitespace": b">c1\tdesc\n  ACGT ACGT\t\n\x0b AC\x0cGT \x1f\n>c2\n A C G T \n",
    "indented_header_and_blank_lines": b"\n\n  >c1  more words\n\nACGT\n\n\n   \nTTTT\n  > c2 name_after_space\nGG\n",
    "duplicate_names": b">a\nAAAA\n>b\nCCCC\n>a\nGGGGGG\n",
    "empty_records": b">a\n>b\n\n>c\nACGT\n>d\n",
    "short_for_trim": b">a\n" + b"ACGT" * 15 + b"\n>b\n" + b"ACGT" * 16 + b"\n>c\nAC\n",
    "long_lines": b">big\n" + b"acgtn" * 20000 + b"\n" + b"ACGT" * 5000 + b"\n",
    "non_ascii_falls_back": ">c1\nACGTé\n>c2 ß\nacgt\n".encode("utf-8"),
    "empty_file": b"",
    "only_junk": b"ACGT\nACGT\n",
}


def _both(path, trim):
    import bwt_algorithm_b200  # noqa: F401
    from bwt_algorithm_b200.pipeline import TandemRepeatFinder

    a = TandemRepeatFinder(path, flank_trim=trim)
    fast = a.load_reference()
    b = TandemRepeatFinder(path, flank_trim=trim)
    slow = b._load_reference_lines()
    return a, fast, b, slow


@pytest.mark.parametrize("trim", [30, 0, 3])
@pytest.mark.parametrize("name", sorted(ADVERSARIAL))
def test_native_parser_equals_line_loop(tmp_path, name, trim):
    path = tmp_path / f"{name}.fa"
    path.write_bytes(ADVERSARIAL[name])
    a, fast, b, slow = _both(str(path), trim)
    assert fast == slow and list(fast) == list(slow)
    assert a.full_sequences == b.full_sequences and a.trim_offsets == b.trim_offsets
    native = a._load_reference_native()
    if name == "non_ascii_falls_back":
        assert native is None
    else:
        assert native == slow


@pytest.mark.parametrize("fa", sorted(glob.glob(os.path.join(CLI, "*.fa*"))), ids=os.path.basename)
def test_fixture_fastas(fa):
    for trim in (30, 0):
        a, fast, b, slow = _both(fa, trim)
        assert a._load_reference_native() == slow and fast == slow
        assert a.full_sequences == b.full_sequences and a.trim_offsets == b.trim_offsets


def test_header_without_name_behaves_like_the_reference(tmp_path):
    path = tmp_path / "bad.fa"
    path.write_bytes(b">\nACGT\n")
    import bwt_algorithm_b200  # noqa: F401
    from bwt_algorithm_b200.pipeline import TandemRepeatFinder

    with pytest.raises(IndexError):            # bwt.py:3735: line[1:].split()[0]
        TandemRepeatFinder(str(path)).load_reference()


@pytest.mark.skipif(not os.path.exists("/root/reference/bwt.py"), reason="reference tree only exists in the build container")
@pytest.mark.parametrize("name", sorted(k for k in ADVERSARIAL if k != "empty_file"))
def test_against_the_reference_loader_itself(tmp_path, name):
    import importlib.util
    import sys

    spec = importlib.util.spec_from_file_location("ref_bwt_fasta", "/root/reference/bwt.py")
    ref = importlib.util.module_from_spec(spec)
    sys.modules["ref_bwt_fasta"] = ref
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache_golden")
    spec.loader.exec_module(ref)
    path = tmp_path / f"{name}.fa"
    path.write_bytes(ADVERSARIAL[name])
    r = ref.TandemRepeatFinder(str(path), flank_trim=3)
    want = r.load_reference()
    import bwt_algorithm_b200  # noqa: F401
    from bwt_algorithm_b200.pipeline import TandemRepeatFinder

    a = TandemRepeatFinder(str(path), flank_trim=3)
    got = a.load_reference()
    assert got == want and list(got) == list(want)
    assert a.full_sequences == r.full_sequences and a.trim_offsets == r.trim_offsets
