"""Host shim (mash_b200/host): .msh Cap'n Proto writer/reader, `mash info` JSON dump, `mash paste`, FASTA/FASTQ reader.
CPU only -- nothing here touches the GPU (the binary links libmashgpu.so but only creates a context when it sketches or
compares)."""
import json
import os
import subprocess

import numpy as np
import pytest

import msh_reader
from fixtures import GOLDEN, read_fastx

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MASH = os.path.join(ROOT, "mash_b200", "host", "mash")


@pytest.fixture(scope="module")
def mash():
    if not os.path.exists(MASH):
        subprocess.check_call(["make", "-C", os.path.dirname(MASH)], stdout=subprocess.DEVNULL)
    return MASH


def run(mash, *args, cwd=None):
    return subprocess.run([mash, *args], cwd=cwd, capture_output=True, text=True, check=True)


def test_golden_json_roundtrips_through_msh(mash, tmp_path):
    # reference test/ref/genomes.json -> .msh (our writer) -> `mash info -d` (our reader + writeJson mirror) == byte identical
    for name in ("ref_genomes.json", "ref_reads.json"):
        src = os.path.join(GOLDEN, name)
        msh = str(tmp_path / (name + ".msh"))
        run(mash, "import-json", src, msh)
        out = run(mash, "info", "-d", msh).stdout
        assert out == open(src).read()


def test_msh_layout_matches_capnp_builder_prediction(mash, tmp_path):
    # genomes.msh written by MallocMessageBuilder has 3 segments; seed 42 -> referenceListOld
    msh = str(tmp_path / "g.msh")
    run(mash, "import-json", os.path.join(GOLDEN, "ref_genomes.json"), msh)
    d = msh_reader.read_msh(msh)
    assert d["segments"] == [77, 1001, 2002]
    assert d["list"] == "referenceListOld" and d["hasLocusList"]
    assert (d["kmer"], d["sketchSize"], d["hashSeed"], d["alphabet"], d["concatenated"], d["noncanonical"], d["preserveCase"]) == (21, 1000, 42, "ACGT", True, False, False)
    gold = json.load(open(os.path.join(GOLDEN, "ref_genomes.json")))
    assert len(d["references"]) == 3
    for r, g in zip(d["references"], gold["sketches"]):
        assert (r["name"], r["comment"], r["length"]) == (g["name"], g["comment"], g["length"])
        assert r["hashes64"] == g["hashes"] and r["hashes32"] is None and r["counts32"] is None and not r["counts32Sorted"]


def test_msh_variants_decode_with_independent_reader(mash, tmp_path):
    # non-default seed -> referenceList slot, hashSeed stored xor 42; 32-bit hashes (k <= 16); many references (far pointers)
    dump = {"kmer": 16, "alphabet": "ACGT", "preserveCase": False, "canonical": True, "sketchSize": 40, "hashType": "MurmurHash3_x64_128",
            "hashBits": 32, "hashSeed": 7, "sketches": []}
    rng = np.random.Generator(np.random.PCG64(3))
    for i in range(300):
        n = int(rng.integers(0, 41))
        hs = sorted(set(int(x) for x in rng.integers(0, 2**32, n)))
        dump["sketches"].append({"name": f"seq{i}", "length": int(rng.integers(1, 2**40)), "comment": "c" * int(rng.integers(0, 30)), "hashes": hs})
    # write in the reference's dump format (CommandInfo.cpp:222-299) so that import-json parses it
    lines = ["{", f'\t"kmer" : {dump["kmer"]},', f'\t"alphabet" : "{dump["alphabet"]}",', '\t"preserveCase" : false,', '\t"canonical" : true,',
             f'\t"sketchSize" : {dump["sketchSize"]},', '\t"hashType" : "MurmurHash3_x64_128",', '\t"hashBits" : 32,', f'\t"hashSeed" : {dump["hashSeed"]},',
             ' \t"sketches" :', "\t["]
    for i, s in enumerate(dump["sketches"]):
        lines += ["\t\t{", f'\t\t\t"name" : "{s["name"]}",', f'\t\t\t"length" : {s["length"]},', f'\t\t\t"comment" : "{s["comment"]}",', '\t\t\t"hashes" :', "\t\t\t["]
        lines += [f"\t\t\t\t{h}" + ("," if j < len(s["hashes"]) - 1 else "") for j, h in enumerate(s["hashes"])]
        lines += ["\t\t\t]", "\t\t}," if i < len(dump["sketches"]) - 1 else "\t\t}"]
    lines += ["\t]", "}"]
    src = tmp_path / "d.json"
    src.write_text("\n".join(lines) + "\n")
    msh = str(tmp_path / "d.msh")
    run(mash, "import-json", str(src), msh)
    d = msh_reader.read_msh(msh)
    assert d["list"] == "referenceList" and d["hashSeed"] == 7 and d["kmer"] == 16 and len(d["segments"]) >= 2
    for r, g in zip(d["references"], dump["sketches"]):
        assert (r["name"], r["comment"], r["length"]) == (g["name"], g["comment"], g["length"])
        assert (r["hashes32"] or []) == g["hashes"] and r["hashes64"] is None
    assert run(mash, "info", "-d", msh).stdout == src.read_text()


def test_paste_concatenates_in_order(mash, tmp_path):
    a, b = str(tmp_path / "a.msh"), str(tmp_path / "b.msh")
    run(mash, "import-json", os.path.join(GOLDEN, "ref_genomes.json"), a)
    run(mash, "import-json", os.path.join(GOLDEN, "ref_genomes.json"), b)
    run(mash, "paste", str(tmp_path / "ab"), a, b)
    d = msh_reader.read_msh(str(tmp_path / "ab.msh"))
    assert [r["name"] for r in d["references"]] == ["genome1.fna", "genome2.fna", "genome3.fna"] * 2
    lines = run(mash, "info", "-t", str(tmp_path / "ab.msh")).stdout.splitlines()
    assert len(lines) == 6 and lines[0].split("\t")[:3] == ["1000", "4639675", "genome1.fna"]


def test_sketch_without_gpu_fails_loudly(mash, tmp_path):
    import mash_b200
    if mash_b200.load_library().mashgpu_device_count() > 0:
        pytest.skip("a GPU is present")
    fa = tmp_path / "x.fa"
    fa.write_text(">s1 comment\nACGTACGTACGTACGTACGTACGTACGT\n")
    p = subprocess.run([mash, "sketch", str(fa)], capture_output=True, text=True)
    assert p.returncode == 1 and "no CUDA device" in p.stderr


def test_malformed_msh_is_rejected_not_read_out_of_bounds(mash, tmp_path):
    # users download .msh files: list shapes are validated the way libcapnp bounds-checks (element size code of hash / count
    # lists, list length vs the hashes read, composite list word count) -- a damaged file ends in an error message, not in
    # reads past the mapping
    import struct
    good = tmp_path / "g.msh"
    run(mash, "import-json", os.path.join(GOLDEN, "ref_genomes.json"), str(good))
    data = bytearray(good.read_bytes())
    nseg = struct.unpack_from("<I", data, 0)[0] + 1
    table = (4 + 4 * nseg + 7) & ~7
    seg_sizes = [struct.unpack_from("<I", data, 4 + 4 * i)[0] for i in range(nseg)]
    # the hash lists of genomes.msh live in segments 1 and 2 behind far pointers: the first word of segment 1 is the landing pad,
    # a list pointer to 1000 64-bit elements
    off = table + 8 * seg_sizes[0]
    w = struct.unpack_from("<Q", data, off)[0]
    assert (w & 3) == 1 and ((w >> 32) & 7) == 5 and (w >> 35) == 1000
    variants = {"declared_32_bit": (w & ~(7 << 32)) | (4 << 32),                       # hashes64 declared as a 32-bit list
                "count_beyond_segment": (w & ((1 << 35) - 1)) | ((2 ** 28) << 35),      # more elements than the segment holds
                "short_list": (w & ((1 << 35) - 1)) | (10 << 35)}                        # fewer hashes than sketchSize: legal, just a short sketch
    for name, word in variants.items():
        d = bytearray(data)
        struct.pack_into("<Q", d, off, word)
        f = tmp_path / f"{name}.msh"
        f.write_bytes(bytes(d))
        p = subprocess.run([mash, "info", "-d", str(f)], capture_output=True)
        err = p.stderr.decode(errors="replace")
        if name == "short_list":
            assert p.returncode == 0, err
        else:
            assert p.returncode == 1 and "not a valid sketch file" in err, (name, p.returncode, err)


def test_sequence_buffer_pool(tmp_path):
    # host/seqbuf.hpp: records parsed by `mash sketch` live in recycled pool buffers (fresh-memory page faults bounded the parse stage);
    # contents through growth and moves, heap for short records, slab reuse, six threads at once
    exe = str(tmp_path / "seqbuf_test")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I" + os.path.join(ROOT, "mash_b200", "host"), os.path.join(ROOT, "tools", "seqbuf_test.cpp"), "-o", exe, "-pthread"])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "bad 0" in out.stdout, out.stdout[-2000:]


@pytest.mark.parametrize("threads,rows,cols", [(1, 0, 5), (1, 1, 1), (3, 7, 3), (1, 1500, 400), (5, 1500, 400), (4, 3, 150000), (8, 200000, 2)])
def test_pair_grid_writer_prints_what_the_stream_operators_print(tmp_path_factory, threads, rows, cols):
    # `mash dist` / `mash triangle` format their rows into memory on the -p threads (host/fastout.hpp) instead of one
    # `cout << ... << endl` per pair; the bytes must be the reference's -- doubles as `ostream << double` prints them (6 significant
    # digits, %g), incl. denormals, rounding boundaries, 0, 1, 1e22 -- and keep their place between other cout output
    exe = str(tmp_path_factory.getbasetemp() / "fastout_test")
    if not os.path.exists(exe):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-I" + os.path.join(ROOT, "mash_b200", "host"), os.path.join(ROOT, "tools", "fastout_test.cpp"), "-o", exe, "-pthread"])
    want = subprocess.run([exe, "stream", str(threads), str(rows), str(cols)], capture_output=True, check=True).stdout
    got = subprocess.run([exe, "fast", str(threads), str(rows), str(cols)], capture_output=True, check=True).stdout
    assert got == want and len(want) > 10
