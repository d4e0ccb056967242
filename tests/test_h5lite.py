"""CPU: the pure-NumPy HDF5 subset (utils/h5lite.py) and the Keras .h5 model layout on top of it (utils/keras_h5.py).
No h5py / TensorFlow exists in this image, so these are round trips and structural checks against the HDF5 file format
specification, not a comparison with real Keras output (DESIGN.md says so)."""
import json
import struct

import numpy as np
import pytest

from alphasnake_zero_b200.utils import h5lite, keras_h5
from alphasnake_zero_b200.utils.alpha_nnet import flatten_weights, init_weights, load_weights


def test_h5lite_round_trip_groups_datasets_attributes(tmp_path):
    rng = np.random.default_rng(0)
    root = h5lite.Group()
    root.attrs["title"] = b"hello"
    root.attrs["numbers"] = np.arange(5, dtype=np.int32)
    root.attrs["names"] = np.array([b"a", b"bcd", b"ef"])
    g = root.group("grp")
    a = rng.standard_normal((3, 4, 5)).astype(np.float32)
    g.dataset("x:0", a).attrs["unit"] = "m"
    g.group("nested").dataset("y", np.arange(7, dtype=np.float64))
    root.dataset("empty", np.zeros((0,), np.float32))
    many = root.group("many")
    for i in range(50):                                           # more members than libhdf5's default leaf node holds
        many.dataset("d%02d" % i, np.full((2,), i, np.float32))
    path = str(tmp_path / "t.h5")
    data = h5lite.write(path, root)
    assert data[:8] == h5lite.SIG and struct.unpack_from("<Q", data, 40)[0] == len(data)      # signature, end-of-file address
    r = h5lite.read(path)
    assert r.attrs["title"] == b"hello" and np.array_equal(r.attrs["numbers"], np.arange(5)) and r.attrs["names"].tolist() == [b"a", b"bcd", b"ef"]
    assert np.array_equal(r["grp/x:0"].data.view(np.uint32), a.view(np.uint32)) and r["grp/x:0"].attrs["unit"] == b"m"
    assert np.array_equal(r["grp/nested/y"].data, np.arange(7.0)) and r["empty"].data.shape == (0,)
    assert sorted(r["many"].keys()) == ["d%02d" % i for i in range(50)] and r["many/d37"].data.tolist() == [37.0, 37.0]
    assert "nope" not in r and "grp/nested" in r
    # every object header / heap / tree / node starts on an 8-byte boundary with the right signature
    assert data.count(b"TREE") >= 4 and data.count(b"SNOD") >= 4 and data.count(b"HEAP") >= 4


def test_reader_follows_continuation_blocks_and_multi_node_trees():
    """structures h5py produces but the writer does not: an object header continued in a second block (attributes added after
    creation) and a group whose B-tree has several symbol-table nodes; built by hand from the format specification"""
    buf = bytearray(96)

    def alloc(b):
        while len(buf) % 8:
            buf.append(0)
        off = len(buf)
        buf.extend(b)
        return off

    def msg(t, body):
        body = body + bytes((-len(body)) % 8)
        return struct.pack("<HHB3x", t, len(body), 0) + body
    # two datasets (scalars, compact layout version 3: the value sits in the header)
    def scalar_ds(v):
        dt = struct.pack("<BBBBI", 0x11, 0x20, 31, 0, 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
        body = msg(1, struct.pack("<BBB5x", 1, 0, 0)) + msg(3, dt) + msg(8, struct.pack("<BBH", 3, 0, 4) + struct.pack("<f", v))
        return alloc(struct.pack("<BxHII4x", 1, 3, 1, len(body)) + body)
    d1, d2 = scalar_ds(1.5), scalar_ds(-2.0)
    heap_data = bytearray(8) + b"aa" + bytes(6) + b"zz" + bytes(6)
    hd = alloc(bytes(heap_data))
    heap = alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), 1, hd))

    def snod(entries):
        b = bytearray(b"SNOD" + struct.pack("<BxH", 1, len(entries)))
        for off, addr in entries:
            b += struct.pack("<QQII16x", off, addr, 0, 0)
        return alloc(bytes(b) + bytes(8 + 40 * 8 - len(b)))
    s1, s2 = snod([(8, d1)]), snod([(16, d2)])
    bt = bytearray(b"TREE" + struct.pack("<BBHQQ", 0, 0, 2, h5lite.UNDEF, h5lite.UNDEF))
    bt += struct.pack("<QQQQQ", 0, s1, 8, s2, 16)
    btree = alloc(bytes(bt) + bytes(544 - len(bt)))
    # root header: symbol table message + a continuation message that points at a block with one attribute
    attr = h5lite._attr_message("late", b"added later")
    cont = alloc(attr)
    body = msg(0x11, struct.pack("<QQ", btree, heap)) + msg(0x10, struct.pack("<QQ", cont, len(attr)))
    root = alloc(struct.pack("<BxHII4x", 1, 3, 1, len(body)) + body)
    sb = h5lite.SIG + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0) + struct.pack("<HHI", 4, 16, 0)
    sb += struct.pack("<QQQQ", 0, h5lite.UNDEF, len(buf), h5lite.UNDEF) + struct.pack("<QQII", 0, root, 1, 0) + struct.pack("<QQ", btree, heap)
    buf[:96] = sb
    r = h5lite.read(bytes(buf))
    assert r.keys() == ["aa", "zz"] and float(r["aa"].data) == 1.5 and float(r["zz"].data) == -2.0
    assert r.attrs["late"] == b"added later"


def test_keras_h5_round_trip_and_layout(tmp_path):
    w = init_weights((21, 21, 3), seed=5)
    w["bn0"]["mean"] = np.linspace(-1, 1, 128).astype(np.float32)            # make every array distinguishable
    path = str(tmp_path / "AlphaSnake7.h5")
    keras_h5.save(w, path)
    back = keras_h5.load(path)
    assert back["side"] == 11
    for a, b in zip(flatten_weights(w), flatten_weights(back)):
        assert a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert all(np.array_equal(a, b) for a, b in zip(flatten_weights(w), flatten_weights(load_weights(path))))
    # the layout tf.keras 2.2.4 writes: model_config + model_weights/<layer>/<layer>/<weight>:0, layer_names / weight_names attributes
    r = h5lite.read(path)
    cfg = json.loads(r.attrs["model_config"].decode())
    assert cfg["class_name"] == "Model" and len(cfg["config"]["layers"]) == 40
    kinds = [l["class_name"] for l in cfg["config"]["layers"]]
    assert kinds.count("Conv2D") == 10 and kinds.count("BatchNormalization") == 10 and kinds.count("Dense") == 2 and kinds.count("Add") == 4
    assert cfg["config"]["layers"][0]["config"]["batch_input_shape"] == [None, 21, 21, 3]
    names = [n.decode() for n in r["model_weights"].attrs["layer_names"]]
    assert names[:4] == ["input_1", "conv2d", "batch_normalization", "activation"] and names[-1] == "activation_11"
    assert r["model_weights/conv2d_3/conv2d_3/kernel:0"].data.shape == (3, 3, 128, 128)
    assert [n.decode() for n in r["model_weights/batch_normalization"].attrs["weight_names"]] == [
        "batch_normalization/gamma:0", "batch_normalization/beta:0", "batch_normalization/moving_mean:0",
        "batch_normalization/moving_variance:0"]
    assert r["model_weights/dense/dense/bias:0"].data.shape == (128,) and r["model_weights/dense_1/dense_1/kernel:0"].data.shape == (128, 3)
    assert np.size(r["model_weights/add_2"].attrs["weight_names"]) == 0


def test_keras_h5_loader_does_not_depend_on_layer_name_counters(tmp_path):
    """a model built later in a session carries arbitrary suffixes (conv2d_37 ...): the loader follows layer_names / weight_names"""
    w = init_weights((13, 13, 3), seed=6)                                     # a 7x7 board
    path = str(tmp_path / "m.h5")
    keras_h5.save(w, path)
    r = h5lite.read(path)
    root = h5lite.Group()
    root.attrs.update(r.attrs)
    mw = root.group("model_weights")
    ren = {}
    for nm in [n.decode() for n in r["model_weights"].attrs["layer_names"]]:
        base = nm.rsplit("_", 1)[0] if nm.rsplit("_", 1)[-1].isdigit() else nm
        idx = int(nm.rsplit("_", 1)[1]) if nm.rsplit("_", 1)[-1].isdigit() else 0
        ren[nm] = "%s_%d" % (base, idx + 40)
    mw.attrs["layer_names"] = np.array([ren[n.decode()].encode() for n in r["model_weights"].attrs["layer_names"]])
    for old, new in ren.items():
        g = mw.group(new)
        src = r["model_weights/" + old]
        wn = [n.decode() for n in np.atleast_1d(src.attrs["weight_names"])] if np.size(src.attrs["weight_names"]) else []
        g.attrs["weight_names"] = np.array([n.replace(old + "/", new + "/").encode() for n in wn]) if wn else np.zeros((0,), "S1")
        if wn:
            inner = g.group(new)
            for n in wn:
                inner.dataset(n.split("/")[1], src[n].data)
    path2 = str(tmp_path / "renamed.h5")
    h5lite.write(path2, root)
    back = keras_h5.load(path2)
    assert back["side"] == 7
    assert all(np.array_equal(a, b) for a, b in zip(flatten_weights(w), flatten_weights(back)))


def test_chunked_layout_is_refused():
    r = h5lite._Reader.__new__(h5lite._Reader)
    r.b = bytes([3, 2]) + bytes(30)
    r.base = 0
    with pytest.raises(NotImplementedError):
        r.parse_layout(0)
