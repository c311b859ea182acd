"""bipymc_b200/h5lite.py -- the package's own HDF5 writer / reader for the reference's checkpoint layout
(bipymc/chain.py:59-93, bipymc/demc.py:198-233).  No libhdf5 exists in the build container, so the on-disk
format is pinned here structure by structure against the HDF5 File Format Specification (version 1.1
structures, what h5py writes with libver="earliest"); tests/test_hdf5_optional.py cross-reads with h5py
wherever it imports."""
import struct
import zlib

import numpy as np
import pytest

from bipymc_b200 import h5lite as h5
from bipymc_b200.chain import McmcChain

UNDEF = 0xFFFFFFFFFFFFFFFF


def _one_chain_file(tmp_path, data):
    f = str(tmp_path / "one.h5")
    with h5.File(f, "w") as h:
        h.create_dataset("/chains/chain_id_7", data=data, compression="gzip")
        h["/chains"].attrs["b200_seed"] = np.uint64(31)
    return f, open(f, "rb").read()


def _messages(raw, addr):
    """(type, flags, body) of the version-1 object header at addr."""
    ver, _, nmsg, refs, size = struct.unpack_from("<BBHII", raw, addr)
    assert ver == 1 and refs == 1
    out, p = [], addr + 16
    while len(out) < nmsg:
        t, n, fl = struct.unpack_from("<HHB", raw, p)
        assert n % 8 == 0
        out.append((t, fl, raw[p + 8:p + 8 + n]))
        p += 8 + n
    assert p == addr + 16 + size
    return out


def test_superblock_and_root_group_bytes(tmp_path):
    f, raw = _one_chain_file(tmp_path, np.arange(12.0).reshape(6, 2))
    # superblock version 0 (spec III.A): signature, versions, sizes of offsets / lengths, K values, addresses
    assert raw[:8] == b"\x89HDF\r\n\x1a\n"
    assert raw[8:16] == bytes([0, 0, 0, 0, 0, 8, 8, 0])
    leaf_k, internal_k, flags = struct.unpack_from("<HHI", raw, 16)
    assert (leaf_k, internal_k, flags) == (4, 16, 0)
    base, free, eof, driver = struct.unpack_from("<QQQQ", raw, 24)
    assert base == 0 and free == UNDEF and driver == UNDEF and eof == len(raw)
    # root symbol-table entry: name offset 0, object header, cache type 1 with B-tree / heap addresses
    name_off, root_oh, cache, _, bt, hp = struct.unpack_from("<QQIIQQ", raw, 56)
    assert name_off == 0 and cache == 1
    msgs = _messages(raw, root_oh)
    assert [m[0] for m in msgs] == [0x11]                       # symbol-table message
    assert struct.unpack("<QQ", msgs[0][2]) == (bt, hp)
    # local heap: "HEAP", version 0, data segment with "" at offset 0, free list that terminates with 1
    assert raw[hp:hp + 4] == b"HEAP" and raw[hp + 4] == 0
    seg_size, free_head, seg = struct.unpack_from("<QQQ", raw, hp + 8)
    assert raw[seg:seg + 8] == b"\0" * 8 and raw[seg + 8:seg + 15] == b"chains\0"
    nxt, fsize = struct.unpack_from("<QQ", raw, seg + free_head)
    assert nxt == 1 and free_head + fsize == seg_size
    # group B-tree: "TREE", type 0, level 0, one child; keys are heap offsets ("" then the largest name)
    assert raw[bt:bt + 4] == b"TREE" and raw[bt + 4] == 0 and raw[bt + 5] == 0
    used, left, right = struct.unpack_from("<HQQ", raw, bt + 6)
    assert used == 1 and left == UNDEF and right == UNDEF
    k0, snod, k1 = struct.unpack_from("<QQQ", raw, bt + 24)
    assert k0 == 0 and k1 == 8
    assert len(raw) >= bt + 24 + 33 * 8 + 32 * 8                # the node is there at its full size (2 K = 32)
    # symbol-table node: "SNOD", version 1, one entry naming the sub-group with its cached B-tree / heap
    assert raw[snod:snod + 4] == b"SNOD" and raw[snod + 4] == 1
    assert struct.unpack_from("<H", raw, snod + 6)[0] == 1
    e_name, e_oh, e_cache = struct.unpack_from("<QQI", raw, snod + 8)
    assert e_name == 8 and e_cache == 1
    assert len(raw) >= snod + 8 + 8 * 40
    sub = _messages(raw, e_oh)
    assert [m[0] for m in sub] == [0x11, 0x0C]                  # symbol table + the attribute
    assert struct.unpack_from("<QQ", raw, snod + 8 + 24) == struct.unpack("<QQ", sub[0][2])


def test_dataset_header_messages_bytes(tmp_path):
    data = np.arange(12.0).reshape(6, 2)
    f, raw = _one_chain_file(tmp_path, data)
    with h5.File(f, "r") as h:
        ds = h["/chains/chain_id_7"]
        assert ds.chunks == (6, 2)
    # locate the dataset's object header through the structures: root -> chains -> chain_id_7
    def child(oh, want):
        bt, hp = struct.unpack("<QQ", [m for m in _messages(raw, oh) if m[0] == 0x11][0][2])
        seg = struct.unpack_from("<Q", raw, hp + 24)[0]
        snod = struct.unpack_from("<Q", raw, bt + 32)[0]
        for i in range(struct.unpack_from("<H", raw, snod + 6)[0]):
            no, o = struct.unpack_from("<QQ", raw, snod + 8 + 40 * i)
            if raw[seg + no:raw.index(b"\0", seg + no)] == want:
                return o
        raise KeyError(want)
    oh = child(child(struct.unpack_from("<Q", raw, 64)[0], b"chains"), b"chain_id_7")
    msgs = dict((m[0], m[2]) for m in _messages(raw, oh))
    # dataspace v1: version, rank, flags, 5 reserved bytes, dimensions
    assert msgs[0x01] == bytes([1, 2, 0, 0, 0, 0, 0, 0]) + struct.pack("<QQ", 6, 2)
    # datatype: the canonical IEEE little-endian double (class 1 version 1, bit field 0x20 0x3f 0x00, size 8;
    # bit offset 0, precision 64, exponent at 52 of 11 bits, mantissa at 0 of 52 bits, bias 1023)
    assert msgs[0x03][:20] == bytes.fromhex("11203f0008000000" "00004000340b0034ff030000")
    # fill value v2: allocate incrementally (3), write if set (2), defined, size 0
    assert msgs[0x05] == bytes([2, 3, 2, 1, 0, 0, 0, 0])
    # filter pipeline v1: one filter, id 1 (deflate), 8-byte name, optional, one client value (level 4) + padding
    assert msgs[0x0B] == bytes([1, 1, 0, 0, 0, 0, 0, 0]) + struct.pack("<HHHH", 1, 8, 1, 1) + b"deflate\0" + \
        struct.pack("<II", 4, 0)
    # layout v3, chunked: rank + 1 dimensions, chunk B-tree address, chunk dims then the element size
    lay = msgs[0x08]
    assert lay[:3] == bytes([3, 2, 3])
    bt = struct.unpack_from("<Q", lay, 3)[0]
    assert struct.unpack_from("<III", lay, 11) == (6, 2, 8)
    # chunk B-tree: type 1, one chunk; key = (stored size, filter mask, offsets incl. the element dimension)
    assert raw[bt:bt + 4] == b"TREE" and raw[bt + 4] == 1 and raw[bt + 5] == 0
    assert struct.unpack_from("<H", raw, bt + 6)[0] == 1
    nbytes, mask, o0, o1, o2, addr = struct.unpack_from("<IIQQQQ", raw, bt + 24)
    assert (mask, o0, o1, o2) == (0, 0, 0, 0)
    assert zlib.decompress(raw[addr:addr + nbytes]) == data.tobytes()
    end = struct.unpack_from("<IIQQQ", raw, bt + 24 + 40)
    assert end == (0, 0, 6, 0, 0)                               # the key past the last chunk
    assert len(raw) >= bt + 24 + 65 * 32 + 64 * 8               # full node (2 K = 64 children)


def test_attribute_message_bytes(tmp_path):
    f = str(tmp_path / "a.h5")
    with h5.File(f, "w") as h:
        g = h.create_group("chains")
        g.attrs["seed"] = np.int32(5)
    raw = open(f, "rb").read()
    at = raw.index(b"seed\0")
    body = raw[at - 8:]
    # attribute v1: version, reserved, name size (with the terminator), datatype size, dataspace size, then the
    # three parts each padded to 8 bytes, then the value
    assert struct.unpack_from("<BBHHH", body, 0) == (1, 0, 5, 12, 8)
    assert body[8:16] == b"seed\0\0\0\0"
    assert body[16:28] == bytes.fromhex("1008000004000000" "00002000")      # signed 32-bit little-endian integer
    assert body[32:40] == bytes([1, 0, 0, 0, 0, 0, 0, 0])                     # scalar dataspace
    assert struct.unpack_from("<i", body, 40)[0] == 5


def test_round_trip_many_chains_multi_level_trees(tmp_path):
    """300 datasets in one group: 38 symbol-table nodes under a two-level group B-tree; a dataset of 140
    chunks: a two-level chunk B-tree."""
    f = str(tmp_path / "many.h5")
    rng = np.random.RandomState(0)
    chains = [rng.randn(9 + i % 3, 2) for i in range(300)]
    big = rng.randn(7000, 40)
    with h5.File(f, "w") as h:
        for i, c in enumerate(chains):
            h.create_dataset("/chains/chain_id_%d" % i, data=c, compression="gzip")
        h["/chains"].attrs["b200_p_cr"] = np.array([0.2, 0.3, 0.5])
        h["/chains"].attrs["note"] = "hello"
        h.create_dataset("big", data=big, compression="gzip", chunks=(50, 40), shuffle=True)
        h.create_dataset("plain", data=np.arange(12, dtype=np.int32).reshape(3, 4))
        h.create_dataset("scalar", data=np.float64(3.5))
        h.create_dataset("f4", data=np.arange(5, dtype=np.float32))
    raw = open(f, "rb").read()
    root_bt = struct.unpack_from("<Q", raw, 80)[0]
    assert raw[root_bt + 5] == 0
    with h5.File(f, "r") as h:
        assert h.keys() == ["big", "chains", "f4", "plain", "scalar"]
        assert len(h["chains"]) == 300 and "chain_id_299" in h["chains"] and "chain_id_300" not in h["chains"]
        for i, c in enumerate(chains):
            ds = h["/chains/chain_id_%d" % i]
            assert ds.shape == c.shape and ds.dtype == np.float64 and ds.compression == "gzip"
            assert ds.compression_opts == 4
            assert np.array_equal(ds[:], c)
        at = dict(h["/chains"].attrs.items())
        assert np.array_equal(at["b200_p_cr"], [0.2, 0.3, 0.5]) and at["note"] == b"hello"
        assert np.array_equal(h["big"][:], big) and h["big"].shuffle and h["big"].chunks == (50, 40)
        assert np.array_equal(h["big"][10:20, 3], big[10:20, 3])
        assert np.array_equal(h["plain"][...], np.arange(12).reshape(3, 4)) and h["plain"].compression is None
        assert h["scalar"][()] == 3.5 and h["scalar"].shape == ()
        assert h["f4"].dtype == np.float32 and np.array_equal(h["f4"][:], np.arange(5))
        with pytest.raises(OSError):
            h.create_dataset("x", data=np.zeros(3))
        with pytest.raises(KeyError):
            h["/chains/nope"]
        big_bt = h["big"]._layout[1]
    # the chains group needs a level-1 B-tree (38 symbol-table nodes > 32 children per node), and so does the
    # chunk index of the big dataset (140 chunks > 64 children per node)
    root_snod = struct.unpack_from("<Q", raw, root_bt + 32)[0]
    names = {}
    seg = struct.unpack_from("<Q", raw, struct.unpack_from("<Q", raw, 88)[0] + 24)[0]
    for i in range(struct.unpack_from("<H", raw, root_snod + 6)[0]):
        no, oh, cache, _, sbt, shp = struct.unpack_from("<QQIIQQ", raw, root_snod + 8 + 40 * i)
        names[raw[seg + no:raw.index(b"\0", seg + no)]] = (cache, sbt)
    assert names[b"chains"][0] == 1 and names[b"big"][0] == 0
    chains_bt = names[b"chains"][1]
    assert raw[chains_bt:chains_bt + 4] == b"TREE" and raw[chains_bt + 4] == 0 and raw[chains_bt + 5] == 1
    assert struct.unpack_from("<H", raw, chains_bt + 6)[0] == 2
    assert raw[big_bt:big_bt + 4] == b"TREE" and raw[big_bt + 4] == 1 and raw[big_bt + 5] == 1
    assert struct.unpack_from("<H", raw, big_bt + 6)[0] == 3


def test_append_mode_and_delete(tmp_path):
    f = str(tmp_path / "ap.h5")
    a, b = np.random.RandomState(1).randn(5, 3), np.ones((4, 3))
    with h5.File(f, "w") as h:
        h.create_dataset("/chains/chain_id_0", data=a, compression="gzip")
        h.create_dataset("/chains/chain_id_1", data=a * 2, compression="gzip")
    with h5.File(f, "a") as h:
        del h["/chains/chain_id_0"]
        assert "/chains/chain_id_0" not in h
        h.create_dataset("/chains/chain_id_0", data=b, compression="gzip")
        with pytest.raises(ValueError):
            h.create_dataset("/chains/chain_id_1", data=b)
    with h5.File(f, "r") as h:
        assert np.array_equal(h["/chains/chain_id_0"][:], b)
        assert np.array_equal(h["/chains/chain_id_1"][:], a * 2)
    with pytest.raises(OSError):
        h5.File(str(tmp_path / "missing.h5"), "r")
    open(str(tmp_path / "junk.h5"), "wb").write(b"x" * 200)
    with pytest.raises(OSError):
        h5.File(str(tmp_path / "junk.h5"), "r")


def test_unsupported_content_is_refused(tmp_path):
    with h5.File(str(tmp_path / "u.h5"), "w") as h:
        with pytest.raises(TypeError):
            h.create_dataset("c", data=np.zeros(3, dtype=np.complex128))
        with pytest.raises(TypeError):
            h.attrs["o"] = np.array([object()])
        with pytest.raises(ValueError):
            h.attrs["huge"] = np.zeros(10000)
        assert "huge" not in h.attrs


def test_reader_follows_continuation_blocks_and_v2_dataspace(tmp_path):
    """Structures h5lite never writes but libhdf5 does: an object header whose messages continue in another
    block (message 0x10), NIL messages, and an attribute (version 3) with a version-2 dataspace."""
    f = str(tmp_path / "c.h5")
    with h5.File(f, "w") as h:
        h.create_dataset("d", data=np.arange(4.0))
    raw = bytearray(open(f, "rb").read())
    root_oh = struct.unpack_from("<Q", raw, 64)[0]
    # a second block at the end of the file: a NIL message, then a version-3 attribute "k" = int32 9
    dt = bytes.fromhex("1008000004000000" "00002000")
    attr = struct.pack("<BBHHHB", 3, 0, 2, len(dt), 4, 0) + b"k\0" + dt + bytes([2, 0, 0, 1]) + struct.pack("<i", 9)
    attr += b"\0" * (-len(attr) % 8)
    block = struct.pack("<HHB3x", 0, 8, 0) + b"\0" * 8 + struct.pack("<HHB3x", 0x0C, len(attr), 0) + attr
    block_at = len(raw)
    raw += block
    # rewrite the root header: symbol-table message + continuation message, 4 messages in total
    st = bytes(raw[root_oh + 16:root_oh + 16 + 24])
    cont = struct.pack("<HHB3x", 0x10, 16, 0) + struct.pack("<QQ", block_at, len(block))
    new_oh = len(raw)
    raw += struct.pack("<BBHII4x", 1, 0, 4, 1, len(st) + len(cont)) + st + cont
    struct.pack_into("<Q", raw, 64, new_oh)
    struct.pack_into("<Q", raw, 40, len(raw))
    open(f, "wb").write(bytes(raw))
    with h5.File(f, "r") as h:
        assert h.attrs["k"] == 9 and np.array_equal(h["d"][:], np.arange(4.0))


def test_mcmc_chain_h5_round_trip(tmp_path, monkeypatch):
    """McmcChain.write_chain_h5 / read_chain_h5 (chain.py:59-93) through the package's HDF5 module: by file
    name and by open file object, dataset /chains/chain_id_<global_id>, gzip, (T, dim) float64."""
    monkeypatch.setenv("BIPYMC_B200_H5LITE", "1")
    np.random.seed(0)
    c = McmcChain(np.zeros(3), varepsilon=1e-2, global_id=7)
    for _ in range(5):
        c.append_sample(np.random.randn(3))
    f = str(tmp_path / "one.h5")
    c.write_chain_h5(f)
    with h5.File(f, "r") as h:
        ds = h["/chains/chain_id_7"]
        assert ds.shape == (6, 3) and ds.dtype == np.float64 and ds.compression == "gzip"
        assert np.array_equal(ds[:], c.chain)
    d = McmcChain(np.zeros(3), varepsilon=0.0, global_id=7)
    d.read_chain_h5(f)
    assert np.array_equal(d.chain, c.chain)
    # several chains into one open file, the way DeMcMpi.save_state of the reference does (demc.py:207-214)
    g = str(tmp_path / "all.h5")
    chains = []
    with h5.File(g, "w") as h:
        for i in range(5):
            ch = McmcChain(np.full(2, float(i)), varepsilon=1e-3, global_id=i)
            ch.append_sample(np.random.randn(2))
            ch.write_chain_h5(h)
            ch.write_chain_h5(h)               # rewriting replaces the dataset
            chains.append(ch)
    with h5.File(g, "r") as h:
        assert h["chains"].keys() == ["chain_id_%d" % i for i in range(5)]
        for ch in chains:
            e = McmcChain(np.zeros(2), varepsilon=0.0, global_id=ch.global_id)
            e.read_chain_h5(h)
            assert np.array_equal(e.chain, ch.chain)
    with pytest.raises(RuntimeError):
        c.write_chain_h5(12345)


def test_round_trip_property(tmp_path):
    """Random shapes, dtypes, chunk shapes and filters survive a write / read cycle, alone and after an append."""
    from hypothesis import given, settings, strategies as st, HealthCheck
    counter = [0]

    @settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
    @given(st.lists(st.integers(1, 9), min_size=1, max_size=3), st.sampled_from(["<f8", "<f4", "<i8", "<i4", "<u2", "|u1", ">f8", ">i4"]),
           st.sampled_from([None, "gzip"]), st.booleans(), st.integers(0, 2 ** 31 - 1), st.data())
    def run(shape, dtype, compression, shuffle, seed, data):
        counter[0] += 1
        f = str(tmp_path / ("p%d.h5" % counter[0]))
        rng = np.random.RandomState(seed)
        a = (rng.randn(*shape) * 100).astype(dtype)
        chunks = None
        if compression or shuffle:
            chunks = tuple(data.draw(st.integers(1, n)) for n in shape) if data.draw(st.booleans()) else None
        with h5.File(f, "w") as h:
            h.create_dataset("/g/sub/a", data=a, compression=compression, shuffle=shuffle, chunks=chunks)
            h["/g"].attrs["v"] = a.ravel()[:3]
        with h5.File(f, "a") as h:
            h.create_dataset("/g/b", data=a.T.copy(), compression=compression)
        with h5.File(f, "r") as h:
            got = h["/g/sub/a"]
            assert got.shape == a.shape and got.dtype == a.dtype
            assert np.array_equal(got[...], a) and np.array_equal(h["/g/b"][...], a.T)
            assert np.array_equal(h["/g"].attrs["v"], a.ravel()[:3])
            assert (got.compression == "gzip") == (compression == "gzip") and got.shuffle == bool(shuffle)

    run()
    with h5.File(str(tmp_path / "c.h5"), "w") as h:
        with pytest.raises(ValueError):
            h.create_dataset("x", data=np.zeros((4, 2)), compression="gzip", chunks=(8, 2))
