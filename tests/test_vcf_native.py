"""The native VCF scanner (csrc/vcf_scan.cpp, host code) against the oracle's plain-Python reader
and the reference's fixtures.  CPU only."""
import gzip
import os

import numpy as np
import pytest

from conftest import FIXTURES
from oracle import tiling_port


@pytest.fixture(scope="module", autouse=True)
def _built():
    import __graft_entry__ as g
    g.build()


def _same(path, chrom, tmp_names=None):
    from sstar_b200.utils.utils import _scan_vcf
    header, got = _scan_vcf(path, chrom)
    want, names = tiling_port.read_vcf_text(path, header, chrom)
    assert names == header
    assert list(got) == list(want)
    for c in want:
        assert np.array_equal(got[c]["POS"], want[c]["POS"]) and got[c]["POS"].dtype == np.int32
        assert np.array_equal(got[c]["GT"], want[c]["GT"]) and got[c]["GT"].dtype == np.int8
        assert list(got[c]["REF"]) == list(want[c]["REF"])
        assert list(got[c]["ALT"]) == list(want[c]["ALT"])
    return header, got


@pytest.mark.parametrize("name,chrom", [("test.vcf", "21"), ("test.vcf", None), ("test.0.vcf", "1")])
def test_reference_fixtures(name, chrom):
    header, got = _same(os.path.join(FIXTURES, name), chrom)
    assert len(header) >= 6 and all(len(v["POS"]) > 0 for v in got.values())


def test_gzip_and_absent_chromosome(tmp_path):
    from sstar_b200.utils.utils import _scan_vcf
    src = open(os.path.join(FIXTURES, "test.vcf"), "rb").read()
    gz = tmp_path / "t.vcf.gz"
    with gzip.open(gz, "wb") as fh:
        fh.write(src)
    _same(str(gz), "21")
    header, got = _scan_vcf(str(gz), "22")
    assert got == {} and len(header) == 6
    with pytest.raises(FileNotFoundError):
        _scan_vcf(str(tmp_path / "nope.vcf"), None)
    bad = tmp_path / "bad.vcf"
    bad.write_text("##fileformat=VCFv4.2\n1\t5\t.\tA\tT\t.\t.\t.\tGT\t0|1\n")
    with pytest.raises(ValueError, match="no #CHROM header"):
        _scan_vcf(str(bad), None)


def test_odd_records(tmp_path):
    # missing alleles, '/' and '|', extra FORMAT fields, multi-allelic ALT, haploid calls,
    # multi-digit alleles, CRLF line ends, chromosomes interleaved, no trailing newline
    lines = ["##fileformat=VCFv4.2", "##contig=<ID=1>",
             "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\ts1\ts2\ts3\ts4",
             "1\t10\t.\tA\tT\t.\tPASS\t.\tGT\t0|1\t1/1\t.|.\t./1",
             "2\t7\trs1\tG\tC,T\t50\tPASS\tAC=1\tGT:DP:GQ\t0|2:12:99\t2/1:3:4\t.:.:.\t1:5:6",
             "1\t25\t.\tAC\tA\t.\t.\t.\tGT\t0|0\t0|10\t12|3\t1|.",
             "X\t3\t.\tT\t<DEL>\t.\t.\t.\tGT:AD\t1\t0\t.\t1|1:4,5",
             "1\t26\t.\tC\tG\t.\t.\t.\tGT\t1|1\t1|1\t1|1\t1|1"]
    for ending, tail in (("\n", "\n"), ("\r\n", ""), ("\n", "")):
        p = tmp_path / f"odd{len(ending)}{len(tail)}.vcf"
        p.write_bytes((ending.join(lines) + tail).encode())
        header, got = _same(str(p), None)
        assert header == ["s1", "s2", "s3", "s4"] and list(got) == ["1", "2", "X"]
        assert got["1"]["GT"][0].tolist() == [[0, 1], [1, 1], [-1, -1], [-1, 1]]
        assert got["2"]["GT"][0].tolist() == [[0, 2], [2, 1], [-1, -1], [1, -1]]
        assert got["1"]["GT"][1].tolist() == [[0, 0], [0, 10], [12, 3], [1, -1]]
        assert list(got["2"]["ALT"]) == ["C"] and list(got["X"]["ALT"]) == ["<DEL>"]
        _same(str(p), "1")


def test_large_synthetic_vcf_round_trip(tmp_path):
    # thousands of lines: exercises the threaded line parser; genotypes must survive text and back
    from sstar_b200.synth import synth_chromosome, write_vcf
    from sstar_b200.utils.utils import _scan_vcf
    ref, tgt, pos = synth_chromosome(1_200_000, 20, 15, True, seed=3)
    path = tmp_path / "big.vcf"
    names = write_vcf(str(path), "7", pos, np.concatenate([ref, tgt], axis=1), missing_rate=0.01, seed=1)
    header, got = _scan_vcf(str(path), "7")
    assert header == names and np.array_equal(got["7"]["POS"], pos)
    want, _ = tiling_port.read_vcf_text(str(path), header, "7")
    assert np.array_equal(got["7"]["GT"], want["7"]["GT"])


def test_streaming_blocks_and_sample_subset(tmp_path, monkeypatch):
    """The scanner streams the file block by block and keeps only the listed samples: tiny blocks
    (lines straddle every block edge, some lines are longer than a block) must give the same arrays
    as one big block, and a sample subset the same calls as slicing the full scan
    (allel.read_vcf(samples=...), sstar/utils/utils.py:80)."""
    from sstar_b200.synth import synth_chromosome, write_vcf
    from sstar_b200.utils.utils import _scan_vcf
    ref, tgt, pos = synth_chromosome(400_000, 12, 9, True, seed=5)
    path = tmp_path / "s.vcf"
    names = write_vcf(str(path), "3", pos, np.concatenate([ref, tgt], axis=1), missing_rate=0.02, seed=2)
    # a second chromosome interleaved at the end and in front: skipped before anything is stored
    body = open(path).read().split("\n")
    head, data = [l for l in body if l.startswith("#")], [l for l in body if l and not l.startswith("#")]
    other = [("9" + l[1:]) for l in data[:50]]
    path.write_text("\n".join(head + other[:25] + data + other[25:]) + "\n")
    header, full = _scan_vcf(str(path), "3")
    assert header == names and list(full) == ["3"] and np.array_equal(full["3"]["POS"], pos)
    for block in (64, 257, 4096):
        monkeypatch.setenv("SSTAR_B200_VCF_BLOCK", str(block))
        h2, got = _scan_vcf(str(path), "3")
        assert h2 == header and np.array_equal(got["3"]["POS"], pos) and np.array_equal(got["3"]["GT"], full["3"]["GT"])
        assert list(got["3"]["REF"]) == list(full["3"]["REF"]) and list(got["3"]["ALT"]) == list(full["3"]["ALT"])
        h3, both = _scan_vcf(str(path), None)
        assert list(both) == ["9", "3"] and len(both["9"]["POS"]) == 50
    monkeypatch.delenv("SSTAR_B200_VCF_BLOCK")
    pick = [names[7], names[0], names[13], "not_in_file", names[7]]
    hs, sub = _scan_vcf(str(path), "3", samples=pick)
    assert hs == [names[0], names[7], names[13]]          # header order, each once
    assert np.array_equal(sub["3"]["GT"], full["3"]["GT"][:, [0, 7, 13], :])
    hs, none = _scan_vcf(str(path), "3", samples=["nobody"])
    assert hs == [] and none["3"]["GT"].shape == (len(pos), 0, 2)


def _write_bgzf(path, data: bytes, chunk: int = 60000):
    """bgzip's container (SAM specification 4.1): gzip members of <= 64 KB with a 'BC' extra field holding
    the member's size - 1, closed by an empty member."""
    import struct
    import zlib
    with open(path, "wb") as fh:
        for i in list(range(0, len(data), chunk)) + [len(data)]:
            raw = data[i:i + chunk] if i < len(data) else b""
            co = zlib.compressobj(6, zlib.DEFLATED, -15)
            comp = co.compress(raw) + co.flush()
            fh.write(struct.pack("<BBBBIBBH", 0x1F, 0x8B, 8, 4, 0, 0, 0xFF, 6) + b"BC" +
                     struct.pack("<HH", 2, len(comp) + 25) + comp + struct.pack("<II", zlib.crc32(raw), len(raw)))


def test_bgzf_members_are_inflated_in_parallel_and_checked(tmp_path, monkeypatch):
    """A bgzip'ed VCF (what tabix-indexed files are): the members are located by their 'BC' sizes and inflated
    by the worker threads; same arrays as the plain file, whatever the member / block sizes (lines straddle
    both kinds of edges); a corrupt or truncated member is an error, not silence."""
    from sstar_b200.synth import synth_chromosome, write_vcf
    from sstar_b200.utils.utils import _scan_vcf
    ref, tgt, pos = synth_chromosome(2_000_000, 30, 20, True, seed=11)
    plain = tmp_path / "p.vcf"
    names = write_vcf(str(plain), "5", pos, np.concatenate([ref, tgt], axis=1), missing_rate=0.01, seed=4)
    data = plain.read_bytes()
    header, want = _scan_vcf(str(plain), "5")
    for chunk, block in ((60000, None), (65280, None), (777, None), (60000, 4096), (5000, 257)):
        bz = tmp_path / f"b{chunk}.vcf.gz"
        _write_bgzf(str(bz), data, chunk)
        assert gzip.open(bz, "rb").read() == data  # (a valid multi-member gzip file)
        if block:
            monkeypatch.setenv("SSTAR_B200_VCF_BLOCK", str(block))
        h2, got = _scan_vcf(str(bz), "5")
        monkeypatch.delenv("SSTAR_B200_VCF_BLOCK", raising=False)
        assert h2 == header == names and np.array_equal(got["5"]["POS"], pos)
        assert np.array_equal(got["5"]["GT"], want["5"]["GT"])
        assert list(got["5"]["REF"]) == list(want["5"]["REF"]) and list(got["5"]["ALT"]) == list(want["5"]["ALT"])
    _same(str(tmp_path / "b60000.vcf.gz"), "5")  # and the oracle's independent reader agrees
    # the reference fixture, bgzip'ed
    fx = tmp_path / "t.vcf.gz"
    _write_bgzf(str(fx), open(os.path.join(FIXTURES, "test.vcf"), "rb").read(), 300)
    _same(str(fx), "21")
    # corruption: a flipped byte inside a member's deflate data, and a file cut inside a member
    raw = bytearray((tmp_path / "b60000.vcf.gz").read_bytes())
    raw[len(raw) // 2] ^= 0x5A
    bad = tmp_path / "bad.vcf.gz"
    bad.write_bytes(bytes(raw))
    with pytest.raises(ValueError, match="BGZF"):
        _scan_vcf(str(bad), "5")
    cut = tmp_path / "cut.vcf.gz"
    cut.write_bytes((tmp_path / "b60000.vcf.gz").read_bytes()[:-100])
    with pytest.raises(ValueError, match="BGZF"):
        _scan_vcf(str(cut), "5")


def test_random_call_fields_take_the_same_values_on_both_parser_paths(tmp_path):
    """Seeded fuzz of the sample fields: the in-place decoder of the common call shape and the general
    field parser must agree with the oracle's reader on every mixture -- one- and multi-digit alleles, '.',
    '|' and '/', haploid calls, FORMAT fields behind the GT, any position in the line (first, last, behind a
    long field)."""
    rng = np.random.default_rng(20261020)
    alle = ["0", "1", "2", ".", "7", "10", "12", "127"]
    tails = ["", "", "", ":12", ":3,4:99", ":."]

    def field():
        kind = rng.integers(0, 10)
        if kind < 6:
            return f"{rng.choice(alle[:4])}{rng.choice(['|', '/'])}{rng.choice(alle[:4])}"
        if kind < 8:
            return f"{rng.choice(alle)}{rng.choice(['|', '/'])}{rng.choice(alle)}{rng.choice(tails)}"
        if kind == 8:
            return f"{rng.choice(alle)}{rng.choice(tails)}"  # haploid
        return rng.choice([".", "./.", ".|.", "0|1:", "1/"])

    n = 37
    names = [f"s{i}" for i in range(n)]
    lines = ["##fileformat=VCFv4.2", "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(names)]
    for v in range(400):
        fmt = "GT" if rng.integers(0, 3) else "GT:DP:GQ"
        lines.append("\t".join([str(1 + v % 3), str(10 + 7 * v), ".", "A", "T,G", ".", ".", ".", fmt] +
                               [field() for _ in range(n)]))
    p = tmp_path / "fuzz.vcf"
    for ending in ("\n", "\r\n"):
        p.write_bytes((ending.join(lines) + ending).encode())
        _same(str(p), None)
        _same(str(p), "2")
    # a sample subset walks the same fields but stores only some of them
    from sstar_b200.utils.utils import _scan_vcf
    header, full = _scan_vcf(str(p), "1")
    pick = [names[0], names[5], names[36]]
    hs, sub = _scan_vcf(str(p), "1", samples=pick)
    assert hs == pick and np.array_equal(sub["1"]["GT"], full["1"]["GT"][:, [0, 5, 36], :])


def test_scanner_is_clean_under_address_and_ub_sanitizers(tmp_path):
    """The scanner compiled with -fsanitize=address,undefined (g++) over the fixtures, BGZF files with tiny members,
    a corrupt and a truncated one, an empty file, a file without header, lines that end early -- at three block
    sizes and two thread counts: no report, and every input ends in a result or a clean error code."""
    import shutil
    import subprocess
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    from conftest import REPO
    main = tmp_path / "main.cpp"
    main.write_text(
        '#include "sstar_b200.h"\n#include <cstdio>\n#include <initializer_list>\n'
        "int main(int argc, char **argv) {\n"
        "  for (int i = 1; i < argc; ++i) for (int nt : {1, 4}) {\n"
        "    sstar_b200_vcf *h = nullptr;\n"
        "    const int rc = sstar_b200_vcf_scan(argv[i], nullptr, nt, &h);\n"
        '    std::printf("%d %lld\\n", rc, rc == 0 && sstar_b200_vcf_n_chroms(h) ? (long long)sstar_b200_vcf_n_variants(h, 0) : -1LL);\n'
        "    if (rc == 0) sstar_b200_vcf_free(h);\n"
        "  }\n  return 0;\n}\n")
    exe = tmp_path / "scan_asan"
    build = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer",
                            "-I", os.path.join(REPO, "include"), "-o", str(exe), str(main),
                            os.path.join(REPO, "sstar_b200", "csrc", "vcf_scan.cpp"), "-lz", "-lpthread"],
                           capture_output=True, text=True)
    if build.returncode != 0:
        pytest.skip("g++ cannot link the sanitizers here: " + build.stderr[-300:])
    fx = open(os.path.join(FIXTURES, "test.vcf"), "rb").read()
    files = {"ok.vcf": fx, "empty.vcf": b"", "nohdr.vcf": b"1\t5\t.\tA\tT\t.\t.\t.\tGT\t0|1",
             "short.vcf": b"#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\ta\tb\n1\t5\n1\t6\t.\tA\tT\t.\t.\t.\tGT\t0|\n"
                          b"1\t7\t.\tA\tT\t.\t.\t.\tGT\t0|1\t1|\n1\t8\t.\tA\tT\t.\t.\t.\tGT\t\t\n1\t9\t.\tA\tT\t.\t.\t.\tGT\t0|1\t1"}
    paths = []
    for name, data in files.items():
        (tmp_path / name).write_bytes(data)
        paths.append(str(tmp_path / name))
    _write_bgzf(str(tmp_path / "m.vcf.gz"), fx, 300)
    raw = bytearray((tmp_path / "m.vcf.gz").read_bytes())
    raw[len(raw) // 2] ^= 0x5A
    (tmp_path / "bad.vcf.gz").write_bytes(bytes(raw))
    (tmp_path / "cut.vcf.gz").write_bytes((tmp_path / "m.vcf.gz").read_bytes()[:-57])
    with gzip.open(tmp_path / "std.vcf.gz", "wb") as fh:
        fh.write(fx)
    paths += [str(tmp_path / n) for n in ("m.vcf.gz", "bad.vcf.gz", "cut.vcf.gz", "std.vcf.gz")]
    want = {"ok.vcf": (0, 19), "empty.vcf": (-1, -1), "nohdr.vcf": (-1, -1), "short.vcf": (0, 5), "m.vcf.gz": (0, 19),
            "bad.vcf.gz": (-1, -1), "cut.vcf.gz": (-1, -1), "std.vcf.gz": (0, 19)}
    for block in ("64", "4096", str(32 << 20)):
        run = subprocess.run([str(exe)] + paths, capture_output=True, text=True, timeout=300,
                             env=dict(os.environ, SSTAR_B200_VCF_BLOCK=block, ASAN_OPTIONS="detect_leaks=1"))
        assert run.returncode == 0 and "Sanitizer" not in run.stderr and "runtime error" not in run.stderr, run.stderr[-2000:]
        got = [tuple(int(x) for x in ln.split()) for ln in run.stdout.strip().splitlines()]
        assert got == [want[os.path.basename(p)] for p in paths for _ in (1, 4)]
