"""f-4: the HDF5 reader and the validation dataset built on it.

No h5py / libhdf5 exists in this image (PARITY UNPINNED for the byte format: the fixtures come from `tests/h5_writer.py`, an
independent writer of the published format).  What IS pinned against the reference: `ValidationDataset`'s item layout, checked
against the reference's class extracted from `datasets/dataset.py:169-238` (the module itself imports h5py)."""
import ast
import os

import numpy as np
import pytest
import torch

from h5_writer import write_h5

REF = "/root/reference/src/diffusion_pde/datasets/dataset.py"


def _fields(seed=0, N=3, C=1, H=8, W=6, T=5):
    rng = np.random.default_rng(seed)
    return dict(A=rng.standard_normal((N, C, H, W)).astype(np.float32), U=rng.standard_normal((N, C, H, W, T)),
                labels=rng.random((N, 2)).astype(np.float32), t_steps=np.linspace(0.0, 0.5, T))


@pytest.mark.parametrize("storage", ["contiguous", "chunked", "chunked+deflate", "chunked+shuffle+deflate"])
def test_reader_roundtrip(tmp_path, storage):
    from dynamical_pde_diffusion_b200.hdf5 import H5File

    d = _fields()
    d["ids"] = np.arange(11, dtype=np.int64)
    d["big_endian"] = np.arange(6, dtype=">f4").reshape(2, 3)
    kw = {}
    if storage != "contiguous":
        kw["U"] = dict(chunks=(2, 1, 8, 4, 2), deflate="deflate" in storage, shuffle="shuffle" in storage)
        kw["A"] = dict(chunks=(1, 1, 5, 6), deflate="deflate" in storage)       # ragged edge chunks
    attrs = {"T": 0.5, "dx": np.float64(1 / 7), "dy": np.float32(0.25), "N": np.int64(3), "name": "heat_logt",
             "description": "vlen:2D heat equation, pseudospectral", "Lx": 1.0, "vec": np.arange(3.0)}
    path = tmp_path / "data.h5"
    write_h5(path, d, attrs, dataset_kw=kw, dataset_attrs={"U": {"units": "vlen:K", "scale": 2.5}})
    with H5File(path) as f:
        assert sorted(f.keys()) == sorted(d) and "labels" in f and "missing" not in f
        for k, v in d.items():
            got = f[k][:]
            assert got.shape == v.shape and got.dtype == v.dtype and np.array_equal(got, v), k
        assert f["U"].shape == d["U"].shape and f["U"][1, 0, 2:4, :, -1].shape == (2, 6)
        assert f.attrs["name"] == "heat_logt" and f.attrs["description"] == "2D heat equation, pseudospectral"
        assert f.attrs["T"] == 0.5 and f.attrs["dx"] == 1 / 7 and f.attrs["dy"] == np.float32(0.25) and f.attrs["N"] == 3
        assert np.array_equal(f.attrs["vec"], np.arange(3.0))
        assert f["U"].attrs == {"units": "K", "scale": 2.5}
        with pytest.raises(KeyError):
            f["nope"]


def test_reader_rejects_what_it_does_not_decode(tmp_path):
    from dynamical_pde_diffusion_b200.hdf5 import H5File, H5FormatError

    p = tmp_path / "x.h5"
    p.write_bytes(b"not an hdf5 file at all" * 10)
    with pytest.raises(H5FormatError, match="signature"):
        H5File(p)
    write_h5(p, {"a": np.zeros(3)}, {})
    raw = bytearray(p.read_bytes())
    raw[8] = 9                                                      # unknown superblock version
    p.write_bytes(bytes(raw))
    with pytest.raises(H5FormatError, match="superblock"):
        H5File(p)


def _reference_validation_dataset():
    tree = ast.parse(open(REF).read())
    body = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "ValidationDataset"]
    ns = {"torch": torch, "np": np}
    exec(compile(ast.Module(body=body, type_ignores=[]), REF, "exec"), ns)
    return ns["ValidationDataset"]


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("time_as_label", [False, True])
@pytest.mark.parametrize("include_t0", [False, True])
@pytest.mark.parametrize("with_labels", [True, False, "1d"])
def test_validation_dataset_matches_the_reference_class(tmp_path, time_as_label, include_t0, with_labels):
    from dynamical_pde_diffusion_b200 import datasets as DS

    d = _fields(seed=3, N=4, C=2, H=5, W=7, T=4)
    if with_labels == "1d":
        d["labels"] = d["labels"][:, 0].copy()
    if not with_labels:
        d.pop("labels")
    path = tmp_path / "val.h5"
    write_h5(path, d, {"dx": 0.1})
    Ref = _reference_validation_dataset()
    ref = Ref(d["U"], d["t_steps"], labels=d.get("labels"), time_as_label=time_as_label, include_t0_as_target=include_t0)
    loader = DS.get_validation_dataloader(path, time_as_label, include_t0)
    assert len(loader.dataset) == len(ref)
    for i, batch in enumerate(loader):
        want = ref[i]
        assert batch["A"].shape == (1, 2, 5, 7) and torch.equal(batch["A"][0], want["A"]) and torch.equal(batch["U"][0], want["U"])
        if want["labels"] is None:
            assert batch["labels"] is None
        else:
            assert torch.equal(batch["labels"][0], want["labels"])
    data, t_steps, labels, attrs = DS.read_validation_file(path)
    assert attrs["dx"] == 0.1 and data.shape == d["U"].shape and (labels is None) == (not with_labels)


@pytest.mark.gpu
def test_test_loop_over_an_hdf5_file(tmp_path):
    """The evaluation caller fed from a data file in the reference's layout: file -> reader -> ValidationDataset -> test_loop."""
    import dynamical_pde_diffusion_b200 as dp
    from conftest import load_golden, net_from_golden
    from dynamical_pde_diffusion_b200 import datasets as DS, evaluation as E

    dev = torch.device("cuda:0")
    gold = load_golden("joint_heat.npz")
    net = net_from_golden(gold, 2, 2, device=dev)
    H, W = 16, 12
    rng = np.random.default_rng(1)
    path = tmp_path / "val.h5"
    write_h5(path, dict(U=rng.standard_normal((2, 1, H, W, 3)).astype(np.float32), t_steps=np.array([0.0, 0.1, 0.2]),
                        labels=np.array([0.3, 0.5], np.float32)), {"dx": 1.0 / (H - 1), "T": 0.2})
    loader = DS.get_validation_dataloader(path, time_as_label=True, include_t0_as_target=False)      # labels = [t, alpha]
    _, _, _, attrs = DS.read_validation_file(path)
    smp = dp.JointSampler(net, dev, (H, W), 2, 3, 1, dp.heat_loss2, {"dx": float(attrs["dx"])}, num_steps=3)
    mask_a, mask_u = E.get_masks((H, W), 0.2, 0.2, 0.05, 0.05, generator=torch.Generator().manual_seed(0))
    res = E.test_loop(smp, loader, 20.0, 0.5, 20.0, mask_a=mask_a, mask_u=mask_u, max_num_samples=3, keep_on_device=True)
    assert res["MAE"].shape == (3, 2, H, W) and np.isfinite(res["MAE"]).all() and np.isfinite(E.summarize(res)).all()
