"""Offline converter between the reference's Keras model files (models/<name><n>.h5, alpha_nnet.py:11-12,108-109) and the
.npz weight files of this engine (AlphaNNet.save / AlphaNNet(model_name)).  SURVEY.md 8(f) #3.

It must run where TensorFlow/Keras (and h5py) are installed -- the reference's own environment; they are not part of the
build image of this repository, so this script is NOT exercised by the test-suite here.

  python tools/convert_keras_weights.py to-npz   models/AlphaSnake12.h5 models/AlphaSnake12.npz
  python tools/convert_keras_weights.py to-keras models/AlphaSnake12.npz models/AlphaSnake12.h5

The .npz holds Keras' `get_weights()` list in order (conv kernel HWIO, then BN gamma, beta, moving_mean, moving_variance
for every convolution; dense kernels (in, out) and biases) as arr_0 ... arr_N plus `side` (board side, from the input shape).
"""
import sys

import numpy as np


def to_npz(h5_path, npz_path):
    from tensorflow.keras.models import load_model
    m = load_model(h5_path)
    w = m.get_weights()
    side = (int(m.input_shape[1]) + 1) // 2
    assert len(w) == 9 * 5 + 5 + 4, "unexpected number of weight arrays: %d" % len(w)   # 9 convs + head conv (kernel + 4 BN) + 2 dense
    np.savez(npz_path, side=side, *w)
    print("wrote %s: %d arrays, board side %d" % (npz_path, len(w), side))


def build_keras(side):
    """the architecture of alpha_nnet.py:13-56, restated with the same layer order so that set_weights lines up"""
    from tensorflow.keras import Input, Model
    from tensorflow.keras.layers import Activation, Add, BatchNormalization, Conv2D, Dense, Flatten
    from tensorflow.keras.regularizers import l2
    n = 2 * side - 1
    x = inp = Input(shape=(n, n, 3))
    reg = l2(1e-5)
    x = Activation("relu")(BatchNormalization(axis=3)(Conv2D(128, 3, padding="same", use_bias=False, kernel_regularizer=reg)(x)))
    for _ in range(4):
        sc = x
        x = Activation("relu")(BatchNormalization(axis=3)(Conv2D(128, 3, padding="same", use_bias=False, kernel_regularizer=reg)(x)))
        x = BatchNormalization(axis=3)(Conv2D(128, 3, padding="same", use_bias=False, kernel_regularizer=reg)(x))
        x = Activation("relu")(Add()([x, sc]))
    x = Activation("relu")(BatchNormalization(axis=3)(Conv2D(1, 1, padding="same", use_bias=False, kernel_regularizer=reg)(x)))
    x = Flatten()(x)
    x = Activation("relu")(Dense(128, kernel_regularizer=reg)(x))
    out = Activation("tanh")(Dense(3, kernel_regularizer=reg)(x))
    return Model(inp, out)


def to_keras(npz_path, h5_path):
    z = np.load(npz_path)
    side = int(z["side"])
    w = [z["arr_%d" % i] for i in range(len(z.files) - 1)]
    m = build_keras(side)
    m.set_weights(w)
    m.compile(optimizer="adam", loss="mean_squared_error")
    m.save(h5_path)
    print("wrote %s" % h5_path)


if __name__ == "__main__":
    if len(sys.argv) != 4 or sys.argv[1] not in ("to-npz", "to-keras"):
        sys.exit(__doc__)
    (to_npz if sys.argv[1] == "to-npz" else to_keras)(sys.argv[2], sys.argv[3])
