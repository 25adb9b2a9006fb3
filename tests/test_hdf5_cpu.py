"""hdf5.py (the from-scratch HDF5 subset behind model.save_weights / load_weights, N1) and the Keras architecture JSON.

  * reader pinned on a GENUINE libhdf5-written file: SciPy ships MATLAB v7.3 test files (HDF5 with a 512-byte user block,
    superblock v0, classic groups); the values are known analytically (0, pi/4, ... 2 pi);
  * writer: byte-level checks of the structures the HDF5 specification prescribes for libver='earliest' files
    (superblock v0 fields, end-of-file address, 8-byte aligned object headers, SNOD / TREE / HEAP signatures, symbol
    nodes sorted by name, at most 8 entries each) and round trips of every weight of the three reference networks in
    the Keras save_weights layout (layer_names / weight_names attributes, <layer>/<layer>/kernel:0 datasets);
  * to_json / model_from_json: Keras 2.3 functional-model JSON fields, and the rebuilt graph equals the original.
"""
import importlib
import json
import os
import struct

import numpy as np
import pytest

from conftest import PKG
from helpers import G, K

H = importlib.import_module(PKG + ".hdf5")
M = importlib.import_module(PKG + ".model")


def test_reader_on_a_genuine_libhdf5_file():
    scipy_io = pytest.importorskip("scipy.io")
    path = os.path.join(os.path.dirname(scipy_io.__file__), "matlab", "tests", "data", "testhdf5_7.4_GLNX86.mat")
    if not os.path.exists(path):
        pytest.skip("SciPy's MATLAB v7.3 test file is not installed")
    root = H.read(path)
    assert list(root) == ["testdouble"]
    ds = root["testdouble"]
    assert ds.data.dtype == np.float64 and ds.data.shape == (9, 1)
    assert np.allclose(ds.data.ravel(), np.arange(9) * np.pi / 4, rtol=0, atol=1e-15)     # MATLAB: 0:pi/4:2*pi
    assert ds.attrs["MATLAB_class"] == b"double"


def _keras_layers(gname, hw=32):
    params, _ = K.init_params(gname, (hw, hw, 1), seed=1)
    graph = G.GRAPHS[gname](hw, 1)
    layers = []
    for l in graph.layers:
        layers.append((l.name, [("%s/%s:0" % (l.name, k), params["%s/%s" % (l.name, k)]) for k in l.weights]))
    return layers, params


@pytest.mark.parametrize("gname", ["unet", "unetpp", "classifier"])
def test_keras_weight_file_round_trip_and_structure(gname, tmp_path):
    layers, params = _keras_layers(gname)
    path = str(tmp_path / "w.h5")
    H.save_keras_weights(path, layers)
    blob = open(path, "rb").read()
    # ---- superblock v0 (HDF5 File Format Specification, "Disk Format: Level 0A") --------------------------------
    assert blob[:8] == b"\x89HDF\r\n\x1a\n" and blob[8] == 0 and blob[13] == 8 and blob[14] == 8
    leaf_k, internal_k = struct.unpack_from("<HH", blob, 16)
    assert (leaf_k, internal_k) == (4, 16)
    base, free, eof, drv = struct.unpack_from("<4Q", blob, 24)
    assert base == 0 and free == H.UNDEF and drv == H.UNDEF and eof == len(blob)
    _n, root_hdr, ctype = struct.unpack_from("<QQI", blob, 56)
    assert ctype == 1 and root_hdr % 8 == 0 and blob[root_hdr] == 1               # cached group entry -> v1 object header
    # ---- every symbol node: signature, <= 2K entries, names sorted -----------------------------------------------
    pos, n_snod = blob.find(b"SNOD"), 0
    while pos >= 0:
        ver, _r, nsym = struct.unpack_from("<BBH", blob, pos + 4)
        assert ver == 1 and 0 <= nsym <= 8 and pos % 8 == 0
        n_snod += 1
        pos = blob.find(b"SNOD", pos + 4)
    n_groups = 1 + len(layers) + sum(1 for _, w in layers if w)                  # root + layer groups + nested name groups
    assert n_snod >= n_groups and blob.count(b"TREE") == n_groups and blob.count(b"HEAP") == n_groups
    # ---- contents -----------------------------------------------------------------------------------------------------
    root = H.read(path)
    assert [n.decode() for n in root.attrs["layer_names"]] == [n for n, _ in layers]       # model.layers order, not sorted
    assert root.attrs["backend"] == b"tensorflow" and root.attrs["keras_version"] == b"2.3.1"
    assert list(root) == sorted(n for n, _ in layers)                                      # links come back in name order
    back = H.load_keras_weights(path)
    assert list(back) == [n for n, _ in layers]
    for lname, ws in layers:
        assert list(back[lname]) == [w for w, _ in ws]
        for wname, val in ws:
            got = back[lname][wname]
            assert got.dtype == np.float32 and np.array_equal(got, val), wname
        if ws:
            assert root[lname][lname][ws[0][0].split("/")[1]].data.shape == ws[0][1].shape    # <layer>/<layer>/kernel:0


def test_full_model_file_layout_and_attribute_kinds(tmp_path):
    """ModelCheckpoint without save_weights_only writes the whole model: the weights sit under /model_weights
    (T1H:1044-1047 -> model.load_weights(filepath), T1H:1073); also: integer / float / string / array attributes,
    big-endian and compact-free datasets, an empty group, > 8 links in one group"""
    layers, _ = _keras_layers("classifier")
    root = H.Group()
    mw = root.require_group("model_weights")
    mw.attrs["layer_names"] = np.array([n.encode() for n, _ in layers])
    mw.attrs["backend"], mw.attrs["keras_version"] = b"tensorflow", b"2.3.1"
    for lname, ws in layers:
        g = mw.require_group(lname)
        g.attrs["weight_names"] = np.array([w.encode() for w, _ in ws]) if ws else np.zeros((0,), "S1")
        for wname, val in ws:
            g.require_group(wname.split("/")[0])[wname.split("/")[1]] = H.Dataset(val)
    root.attrs["model_config"] = json.dumps({"class_name": "Sequential"}).encode()
    root.attrs["epoch"] = np.int64(12)
    root.attrs["lr"] = np.float32(0.0005)
    root.attrs["shape"] = np.array([224, 224, 3], np.int32)
    extra = root.require_group("optimizer_weights")
    for k in range(20):
        extra["slot_%02d" % k] = H.Dataset(np.full((3,), k, np.float64), attrs={"index": np.int32(k)})
    root["big_endian"] = H.Dataset(np.arange(5, dtype=">i4"))
    root["empty_group"] = H.Group()
    path = str(tmp_path / "full.hdf5")
    H.write(path, root)
    r = H.read(path)
    assert int(r.attrs["epoch"]) == 12 and float(r.attrs["lr"]) == np.float32(0.0005) and r.attrs["shape"].tolist() == [224, 224, 3]
    assert json.loads(bytes(r.attrs["model_config"]).decode())["class_name"] == "Sequential"
    assert len(r["optimizer_weights"]) == 20 and int(r["optimizer_weights"]["slot_07"].attrs["index"]) == 7
    assert r["optimizer_weights"]["slot_19"].data.tolist() == [19.0, 19.0, 19.0]
    assert r["big_endian"].data.tolist() == [0, 1, 2, 3, 4] and len(r["empty_group"]) == 0
    back = H.load_keras_weights(path)                              # finds /model_weights like keras.load_weights does
    assert list(back) == [n for n, _ in layers] and back["dense_1"]["dense_1/kernel:0"].shape == (1024, 32)


def test_reader_rejects_what_it_does_not_support(tmp_path):
    p = str(tmp_path / "not.h5")
    open(p, "wb").write(b"PK\x03\x04" + bytes(600))
    with pytest.raises(H.H5Error):
        H.read(p)
    blob = bytearray(b"\x89HDF\r\n\x1a\n" + bytes(200))
    blob[8] = 2                                                   # superblock v2 = libver 'latest'
    open(p, "wb").write(bytes(blob))
    with pytest.raises(H.H5Error, match="superblock version 2"):
        H.read(p)


@pytest.mark.parametrize("gname", ["unet", "unetpp", "classifier"])
def test_to_json_is_keras_functional_json_and_round_trips(gname):
    graph = G.GRAPHS[gname](64, 1)
    cfg = M.keras_config(graph)
    text = json.dumps(cfg)
    assert cfg["class_name"] == "Model" and cfg["keras_version"] == "2.3.1" and cfg["backend"] == "tensorflow"
    lay = cfg["config"]["layers"]
    assert lay[0]["class_name"] == "InputLayer" and lay[0]["config"]["batch_input_shape"] == [None, 64, 64, 1]
    conv = next(l for l in lay if l["class_name"] == "Conv2D")
    assert conv["config"]["kernel_initializer"]["class_name"] == "VarianceScaling" and conv["config"]["padding"] == "same"
    assert conv["inbound_nodes"] == [[["input_1", 0, 0, {}]]]
    cat = [l for l in lay if l["class_name"] == "Concatenate"]
    if gname != "classifier":
        assert len(cat[0]["inbound_nodes"][0]) >= 2 and cat[0]["config"]["axis"] == 3
    assert cfg["config"]["output_layers"] == [[graph.output.producer.name, 0, 0]]
    # rebuild: same layers, names, weight specs (weights then map 1:1)
    rebuilt = M.model_from_json(text)
    assert [(l.name, l.kind) for l in rebuilt.graph.layers] == [(l.name, l.kind) for l in graph.layers]
    assert rebuilt.graph.weight_specs() == graph.weight_specs()
    assert json.loads(rebuilt.to_json()) == cfg
