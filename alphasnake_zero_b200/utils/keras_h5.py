"""Keras 2.x `.h5` model files for the value network (code/utils/alpha_nnet.py:11-12 `load_model(model_name)`, :108-109
`v_net.save('models/' + name + '.h5')`; train.py:32-35 loads `models/<name><generation>.h5`), read and written with the pure-NumPy
HDF5 subset of h5lite.py -- no h5py, no TensorFlow.

File structure written by `Model.save(path)` of tf.keras 2.2.4 (TensorFlow 2.1, the reference's pinned version) and reproduced here:

  /                      attrs: keras_version, backend, model_config (JSON of the functional model)
  /model_weights         attrs: layer_names (every layer of model.layers, in order), backend, keras_version
  /model_weights/<layer> attrs: weight_names (e.g. b'conv2d/kernel:0'; empty for layers without weights)
  /model_weights/<layer>/<layer>/kernel:0 ...   float32 datasets; a weight name contains one '/', so the dataset sits in a nested group

Weight order = Keras' `get_weights()`: Conv2D kernel (HWIO); BatchNormalization gamma, beta, moving_mean, moving_variance;
Dense kernel (in, out), bias.  The optimizer state is not written (the reference re-compiles every generation,
alpha_snake_zero_trainer.py:79,83) and ignored when present.

NOT VALIDATED AGAINST REAL KERAS FILES in this image (see h5lite.py); the loader does not depend on exact layer names -- it follows
`layer_names` / `weight_names` when they are present and otherwise orders the layers by kind and numeric suffix."""
import json
import re

import numpy as np

from . import h5lite

K = 128
L2 = 9.999999747378752e-06      # float32(1e-5) as Keras serialises it


def _layer_plan(n):
    """[(class, name, inbound layer names)] of alpha_nnet.py:19-56 with tf.keras' default names in a fresh session"""
    plan = [("InputLayer", "input_1", [])]
    cnt = {"conv2d": 0, "batch_normalization": 0, "activation": 0, "add": 0, "dense": 0}

    def name(kind):
        i = cnt[kind]
        cnt[kind] += 1
        return kind if i == 0 else "%s_%d" % (kind, i)

    def add(cls, kind, inbound):
        nm = name(kind)
        plan.append((cls, nm, inbound))
        return nm
    h = add("Conv2D", "conv2d", ["input_1"])
    h = add("BatchNormalization", "batch_normalization", [h])
    h = add("Activation", "activation", [h])
    for _ in range(4):
        sc = h
        h = add("Conv2D", "conv2d", [h]); h = add("BatchNormalization", "batch_normalization", [h]); h = add("Activation", "activation", [h])
        h = add("Conv2D", "conv2d", [h]); h = add("BatchNormalization", "batch_normalization", [h])
        h = add("Add", "add", [h, sc]); h = add("Activation", "activation", [h])
    h = add("Conv2D", "conv2d", [h]); h = add("BatchNormalization", "batch_normalization", [h]); h = add("Activation", "activation", [h])
    plan.append(("Flatten", "flatten", [h])); h = "flatten"
    h = add("Dense", "dense", [h]); h = add("Activation", "activation", [h])
    h = add("Dense", "dense", [h]); h = add("Activation", "activation", [h])
    return plan


def _model_config(n):
    glorot = {"class_name": "GlorotUniform", "config": {"seed": None}}
    zeros, ones = {"class_name": "Zeros", "config": {}}, {"class_name": "Ones", "config": {}}
    reg = {"class_name": "L1L2", "config": {"l1": 0.0, "l2": L2}}
    plan = _layer_plan(n)
    layers = []
    n_conv = n_act = n_dense = 0
    total_act = sum(1 for c, _, _ in plan if c == "Activation")
    for cls, nm, inbound in plan:
        cfg = {"name": nm, "trainable": True, "dtype": "float32"}
        if cls == "InputLayer":
            cfg = {"batch_input_shape": [None, n, n, 3], "dtype": "float32", "sparse": False, "ragged": False, "name": nm}
        elif cls == "Conv2D":
            head = n_conv == 9
            n_conv += 1
            cfg.update(filters=1 if head else K, kernel_size=[1, 1] if head else [3, 3], strides=[1, 1],
                       padding="valid" if head else "same", data_format="channels_last", dilation_rate=[1, 1], activation="linear",
                       use_bias=False, kernel_initializer=glorot, bias_initializer=zeros, kernel_regularizer=reg, bias_regularizer=None,
                       activity_regularizer=None, kernel_constraint=None, bias_constraint=None)
        elif cls == "BatchNormalization":
            cfg.update(axis=[3], momentum=0.99, epsilon=0.001, center=True, scale=True, beta_initializer=zeros, gamma_initializer=ones,
                       moving_mean_initializer=zeros, moving_variance_initializer=ones, beta_regularizer=None, gamma_regularizer=None,
                       beta_constraint=None, gamma_constraint=None)
        elif cls == "Activation":
            n_act += 1
            cfg.update(activation="tanh" if n_act == total_act else "relu")
        elif cls == "Flatten":
            cfg.update(data_format="channels_last")
        elif cls == "Dense":
            n_dense += 1
            cfg.update(units=K if n_dense == 1 else 3, activation="linear", use_bias=True, kernel_initializer=glorot, bias_initializer=zeros,
                       kernel_regularizer=reg, bias_regularizer=None, activity_regularizer=None, kernel_constraint=None, bias_constraint=None)
        layers.append({"class_name": cls, "config": cfg, "name": nm,
                       "inbound_nodes": [[[i, 0, 0, {}] for i in inbound]] if inbound else []})
    return {"class_name": "Model",
            "config": {"name": "model", "layers": layers, "input_layers": [["input_1", 0, 0]], "output_layers": [[plan[-1][1], 0, 0]]},
            "keras_version": "2.2.4-tf", "backend": "tensorflow"}


def _arrays_by_layer(w):
    """[(class, [(weight suffix, array)])] in the order of _layer_plan's weighted layers"""
    def bn(p):
        return [("gamma:0", p["gamma"]), ("beta:0", p["beta"]), ("moving_mean:0", p["mean"]), ("moving_variance:0", p["var"])]
    out = [[("kernel:0", w["conv0"])], bn(w["bn0"])]
    for b in range(4):
        for j in range(2):
            out += [[("kernel:0", w["res%d_conv%d" % (b, j)])], bn(w["res%d_bn%d" % (b, j)])]
    out += [[("kernel:0", w["head_conv"])], bn(w["head_bn"])]
    out += [[("kernel:0", w["dense1_w"]), ("bias:0", w["dense1_b"])], [("kernel:0", w["dense2_w"]), ("bias:0", w["dense2_b"])]]
    return out


def save(weights, path):
    """alpha_nnet.py:108-109: write `path` (…/<name>.h5) in the layout tf.keras 2.2.4's Model.save produces"""
    n = 2 * int(weights["side"]) - 1
    plan = _layer_plan(n)
    per_layer = iter(_arrays_by_layer(weights))
    root = h5lite.Group()
    root.attrs["keras_version"] = b"2.2.4-tf"
    root.attrs["backend"] = b"tensorflow"
    root.attrs["model_config"] = json.dumps(_model_config(n)).encode("utf8")
    mw = root.group("model_weights")
    mw.attrs["layer_names"] = np.array([nm.encode("utf8") for _, nm, _ in plan])
    mw.attrs["backend"] = b"tensorflow"
    mw.attrs["keras_version"] = b"2.2.4-tf"
    for cls, nm, _ in plan:
        g = mw.group(nm)
        if cls in ("Conv2D", "BatchNormalization", "Dense"):
            arrs = next(per_layer)
            g.attrs["weight_names"] = np.array([("%s/%s" % (nm, suffix)).encode("utf8") for suffix, _ in arrs])
            inner = g.group(nm)
            for suffix, a in arrs:
                inner.dataset(suffix, np.ascontiguousarray(a, dtype=np.float32))
        else:
            g.attrs["weight_names"] = np.zeros((0,), dtype="S1")
    h5lite.write(path, root)


_KIND_ORDER = ("conv2d", "batch_normalization", "dense")


def _suffix_number(name):
    m = re.search(r"_(\d+)$", name)
    return int(m.group(1)) if m else 0


def _str(x):
    return x.decode("utf8") if isinstance(x, (bytes, np.bytes_)) else str(x)


def load(path):
    """alpha_nnet.py:11-12: the weight dictionary of alphasnake_zero_b200.utils.alpha_nnet from a Keras .h5 model (or weights) file"""
    root = h5lite.read(path)
    mw = root["model_weights"] if "model_weights" in root else root
    layers = {}
    order = [_str(x) for x in np.atleast_1d(mw.attrs["layer_names"])] if "layer_names" in mw.attrs else list(mw.keys())
    for nm in order:
        if nm not in mw.children or not mw.children[nm].is_group:
            continue
        g = mw.children[nm]
        if "weight_names" in g.attrs and np.size(g.attrs["weight_names"]):
            arrs = [np.asarray(g[_str(wn)].data, np.float32) for wn in np.atleast_1d(g.attrs["weight_names"])]
        else:
            found = dict(g.visit_datasets())
            if not found:
                continue
            def pick(key):
                hits = [v for k, v in found.items() if k.split("/")[-1].startswith(key)]
                return np.asarray(hits[0], np.float32) if hits else None
            arrs = [a for a in (pick("kernel"), pick("bias"), pick("gamma"), pick("beta"), pick("moving_mean"), pick("moving_variance"))
                    if a is not None]
        if arrs:
            layers[nm] = arrs
    convs = sorted([k for k, v in layers.items() if v[0].ndim == 4], key=_suffix_number)
    bns = sorted([k for k, v in layers.items() if len(v) == 4 and v[0].ndim == 1], key=_suffix_number)
    denses = sorted([k for k, v in layers.items() if v[0].ndim == 2], key=_suffix_number)
    if "layer_names" in mw.attrs:         # the file's own order is authoritative (names may carry arbitrary session counters)
        pos = {nm: i for i, nm in enumerate(order)}
        convs.sort(key=pos.get); bns.sort(key=pos.get); denses.sort(key=pos.get)
    if len(convs) != 10 or len(bns) != 10 or len(denses) != 2:
        raise ValueError("%s: expected 10 convolutions, 10 batch normalisations and 2 dense layers, found %d / %d / %d"
                         % (path, len(convs), len(bns), len(denses)))

    def bn(nm):
        g, b, m, v = layers[nm]
        return dict(gamma=g, beta=b, mean=m, var=v)
    n = int(round(np.sqrt(layers[denses[0]][0].shape[0])))
    w = {"side": (n + 1) // 2}
    w["conv0"] = layers[convs[0]][0]; w["bn0"] = bn(bns[0])
    for b in range(4):
        for j in range(2):
            w["res%d_conv%d" % (b, j)] = layers[convs[1 + 2 * b + j]][0]
            w["res%d_bn%d" % (b, j)] = bn(bns[1 + 2 * b + j])
    w["head_conv"] = layers[convs[9]][0]; w["head_bn"] = bn(bns[9])
    w["dense1_w"], w["dense1_b"] = layers[denses[0]]
    w["dense2_w"], w["dense2_b"] = layers[denses[1]]
    if w["conv0"].shape != (3, 3, 3, K) or w["head_conv"].shape != (1, 1, K, 1) or w["dense2_w"].shape != (K, 3):
        raise ValueError("%s does not hold the network of alpha_nnet.py:19-56" % path)
    return w
