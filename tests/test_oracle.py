"""Pin the numpy fp64 oracle against the golden fixtures produced by the LIVE reference."""
import numpy as np
import pytest

from oracle import v1t_oracle as O
from golden_util import CASES, Golden, rel_err

# the reference computes in fp32; the oracle in fp64 -> agreement limited by fp32 round-off
TOL_FWD = 2e-5
TOL_GRAD = 2e-4


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_reference_golden(case):
    g = Golden(case)
    cfg = g.core_config()
    for mouse_id, d in g.mice.items():
        out = O.path_forward_backward(
            g.sd, cfg, mouse_id, d["images"], d["behaviors"], d["pupil_centers"], d["y_true"],
            ds_size=g.meta["ds_size"], noise=d.get("noise"))
        assert rel_err(out["fmap"], d["fmap"]) < TOL_FWD
        assert rel_err(out["z"], d["z"]) < TOL_FWD
        assert rel_err(out["y"], d["y"]) < TOL_FWD
        assert abs(out["loss"] - float(d["loss"])) / abs(float(d["loss"])) < TOL_FWD
        assert rel_err(out["dimages"], d["dimages"]) < TOL_GRAD
        assert set(d["grads"]) <= set(out["grads"]) | {k for k in d["grads"] if np.abs(d["grads"][k]).max() == 0}
        for k, ref in d["grads"].items():
            if k not in out["grads"]:
                continue
            got = out["grads"][k]
            assert got.shape == ref.shape, k
            if np.abs(ref).max() == 0:
                assert np.abs(got).max() < 1e-12, k
            else:
                assert rel_err(got, ref) < TOL_GRAD, (k, rel_err(got, ref))


def test_grid_sample_zero_padding_and_clamp():
    """Edge cases the reference path has: clamped positions, shifts pushing corners out of bounds."""
    rng = np.random.default_rng(0)
    B, gh, gw, C, N = 2, 5, 7, 3, 6
    fmap = rng.standard_normal((B, gh, gw, C))
    mu = np.array([[-1.0, -1.0], [1.0, 1.0], [0.0, 0.0], [0.97, -0.99], [-3.0, 2.0], [0.3, 0.4]])
    sigma = np.zeros((N, 2, 2))
    feats = rng.standard_normal((C, N))
    shifts = np.array([[0.5, -0.5], [-2.5, 0.1]])
    z, cache = O.readout_forward(fmap, mu, sigma, feats, None, noise=None, shifts=shifts)
    # sample fully outside the map -> exactly zero (padding_mode="zeros")
    assert z[1, 0] == 0.0
    # exact corner hit without shift reproduces the pixel
    z0, _ = O.readout_forward(fmap, mu, sigma, feats, None)
    np.testing.assert_allclose(z0[:, 0], fmap[:, 0, 0, :] @ feats[:, 0], rtol=1e-12)
    np.testing.assert_allclose(z0[:, 1], fmap[:, gh - 1, gw - 1, :] @ feats[:, 1], rtol=1e-12)
    # mu outside [-1,1] is clamped before sampling and gets zero position-gradient
    G, _ = O.readout_backward(cache, sigma, feats, np.ones((B, N)))
    assert np.all(G["mu"][4] == 0.0)


def test_oracle_finite_difference_loss_grad():
    """Independent check of the hand-derived backward: central differences on a few parameters."""
    g = Golden("tiny_train")
    cfg = g.core_config()
    d = g.mice["A"]
    sd = {k: np.asarray(v, dtype=np.float64) for k, v in g.sd.items()}
    run = lambda s, grads: O.path_forward_backward(
        s, cfg, "A", d["images"], d["behaviors"], d["pupil_centers"], d["y_true"], ds_size=4500,
        noise=d["noise"], want_grads=grads)
    base = run(sd, True)
    rng = np.random.default_rng(1)
    for key in ["core.transformer.blocks.0.mha.to_qkv.weight", "core.patch_embedding.pos_embedding",
                "core.transformer.blocks.1.b-mlp.models.share.0.weight", "readouts.A.sigma",
                "readouts.A.mu_transform.0.weight", "core_shifter.A.mlp.2.weight",
                "core.transformer.blocks.1.mlp.model.0.weight"]:
        idx = tuple(rng.integers(0, s) for s in sd[key].shape)
        h = 1e-5
        sp, sm = dict(sd), dict(sd)
        sp[key] = sd[key].copy(); sp[key][idx] += h
        sm[key] = sd[key].copy(); sm[key][idx] -= h
        fd = (run(sp, False)["loss"] - run(sm, False)["loss"]) / (2 * h)
        an = base["grads"][key][idx]
        assert abs(fd - an) <= 1e-5 * max(1.0, abs(an)) + 1e-6 * abs(base["loss"]) * 0 + 1e-4 * abs(an), (key, fd, an)


@pytest.mark.parametrize("case", ["tiny_train", "color_mode4", "nobias_mu_param", "default_dims"])
def test_torch_port_matches_reference_golden(case):
    """The CPU-baseline port (same ATen ops as the reference) reproduces the reference's fp32 outputs."""
    import torch
    from oracle import torch_port as TP

    g = Golden(case)
    cfg = g.core_config()
    for mouse_id, d in g.mice.items():
        sd = {k: torch.from_numpy(np.asarray(v)).clone() for k, v in g.sd.items()}
        for k, v in sd.items():
            if v.is_floating_point() and not k.endswith(("reg_scale", "scale", "keep_prop", "source_grid", "grid", "one")):
                v.requires_grad_(True)
        tt = lambda a: torch.from_numpy(np.asarray(a))
        loss, y = TP.step(sd, cfg, mouse_id, tt(d["images"]), tt(d["behaviors"]), tt(d["pupil_centers"]),
                          tt(d["y_true"]), ds_size=4500, noise=tt(d["noise"]) if "noise" in d else None)
        assert rel_err(y.numpy(), d["y"]) < 1e-5
        assert abs(loss.item() - float(d["loss"])) / abs(float(d["loss"])) < 1e-5
        for k, ref in d["grads"].items():
            if np.abs(ref).max() > 0:
                assert rel_err(sd[k].grad.numpy(), ref) < 1e-4, k
