"""Keras .h5 checkpoint mapping.  h5py is not installed here, so the HDF5 container is replaced by a
stand-in with the same access pattern (attrs / create_group / create_dataset / [] / in); what is checked is the part this
repo owns: layer order, weight order, shapes, and the error paths."""
import numpy as np
import pytest

from othellozero_b200 import net


class _Group(dict):
    def __init__(self):
        super().__init__()
        self.attrs = {}

    def create_group(self, name):
        g = _Group()
        self[name] = g
        return g

    def create_dataset(self, name, data):
        self[name] = np.array(data)

    def __getitem__(self, key):      # h5py resolves 'layer/kernel:0' style paths relative to the group
        return dict.__getitem__(self, key)


def test_h5_roundtrip_and_layer_order():
    n, C = 6, 128
    blob = net.init_weights(n, C, seed=4, randomize_bn=True)
    root = _Group()
    net.write_h5_group(root, blob, n, C)
    names = [x.decode() for x in root.attrs["layer_names"]]
    assert names == ["conv2d", "batch_normalization", "conv2d_1", "batch_normalization_1", "conv2d_2",
                     "batch_normalization_2", "conv2d_3", "batch_normalization_3", "dense", "batch_normalization_4",
                     "dense_1", "batch_normalization_5", "pi", "v"]
    assert [x.decode() for x in root["batch_normalization"].attrs["weight_names"]] == [
        "batch_normalization/gamma:0", "batch_normalization/beta:0", "batch_normalization/moving_mean:0",
        "batch_normalization/moving_variance:0"]
    assert root["conv2d"]["conv2d/kernel:0"].shape == (3, 3, 2, C) and root["dense"]["dense/kernel:0"].shape == (4 * C, 1024)
    back = net.blob_from_h5_group(root, n, C)
    assert back.dtype == np.float32 and np.array_equal(back, blob)
    # a full-model file keeps the same groups under 'model_weights'; weight-less layers are listed with no weights
    full = _Group()
    mw = full.create_group("model_weights")
    net.write_h5_group(mw, blob, n, C)
    mw.attrs["layer_names"] = [b"input_1"] + list(mw.attrs["layer_names"]) + [b"pi-reshaped"]
    for extra in ("input_1", "pi-reshaped"):
        mw.create_group(extra).attrs["weight_names"] = []
    assert np.array_equal(net.blob_from_h5_group(full, n, C), blob)


def test_h5_shape_and_count_errors():
    n, C = 6, 128
    root = _Group()
    net.write_h5_group(root, net.init_weights(n, C, seed=1), n, C)
    with pytest.raises(ValueError, match="shape"):
        net.blob_from_h5_group(root, 8, C)          # fc1 / pi shapes differ for another board size
    root.attrs["layer_names"] = root.attrs["layer_names"][:-1]
    with pytest.raises(ValueError, match="weight arrays"):
        net.blob_from_h5_group(root, n, C)


def test_h5_needs_h5py_and_says_so(tmp_path):
    try:
        import h5py  # noqa: F401
    except ImportError:
        with pytest.raises(RuntimeError, match="h5py"):
            net.load_keras_h5(str(tmp_path / "x.h5"), 6, 128)
        return
    blob = net.init_weights(6, 128, seed=2)
    net.save_keras_h5(str(tmp_path / "x.h5"), blob, 6, 128)
    assert np.array_equal(net.load_keras_h5(str(tmp_path / "x.h5"), 6, 128), blob)
