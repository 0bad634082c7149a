"""CPU: the oracle against the golden vectors produced by the UNMODIFIED reference
(oracle/gen_golden.py, run in the build container).  fp32 on both sides: tolerances are
summation-order noise."""
import os

import numpy as np
import pytest
import torch

from oracle import unet_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    return torch.load(os.path.join(GOLD, name), weights_only=False)


@pytest.mark.parametrize("case", ["attn_bf32", "unet_bf16", "attn_bf16_ds", "attn_bf16_convT"])
def test_model_cases(case):
    g = _load("models.pt")[case]
    cfg = g["cfg"]
    sd = O.synthetic_state_dict(g["seed"], **cfg)
    x, t = O.synthetic_batch(g["n"], g["hw"], g["hw"], seed=g["seed"] + 100, fg_fraction=g["fg_fraction"])
    ocfg = {k: cfg[k] for k in ("bilinear", "deep_supervision")}
    ev = O.unet_forward(x, sd, attention=cfg["attention"], training=False, **ocfg)
    assert torch.allclose(ev, g["eval_logits"], rtol=1e-4, atol=1e-5)
    loss, logits, grads = O.train_grads(x, t, sd, attention=cfg["attention"], loss_scale=g["loss_scale"], **ocfg)
    assert torch.allclose(logits, g["train_logits"], rtol=1e-3, atol=1e-4)
    assert abs(loss.item() - g["loss"]) < 1e-5
    for k, ref in g["grads"].items():
        assert abs(grads[k].norm().item() - ref["norm"]) <= 2e-3 * ref["norm"] + 1e-7, k
        assert torch.allclose(grads[k].flatten()[:8], ref["head"], rtol=2e-2, atol=1e-3 * ref["norm"] + 1e-8), k
    for k, v in g["buffers"].items():
        assert torch.allclose(sd[k].float(), v.float(), rtol=1e-4, atol=1e-6), k


def test_loss_cases():
    g = _load("loss.pt")
    z, t = g["z"], g["t"]
    for name, fn in (("dice_bce", O.dice_bce_loss), ("dice", O.dice_loss),
                     ("balanced_ce", lambda a, b: O.balanced_ce(a, b, 0.3))):
        zr = z.clone().requires_grad_(True)
        v = fn(zr, t)
        v.backward()
        assert abs(v.item() - g[name]["value"]) < 1e-6, name
        assert torch.allclose(zr.grad, g[name]["grad"], rtol=1e-4, atol=1e-9), name


def test_metrics_cases():
    for g in _load("metrics.pt"):
        cm = O.confusion_matrix(g["z"], g["t"], g["c"], g["ignore"])
        assert np.array_equal(cm, g["cm"].numpy())
        res = O.metrics_from_confusion(cm)
        for k, v in g["result"].items():
            assert abs(res[k] - v) < 1e-12
        assert torch.allclose(O.iou_per_class(g["z"], g["t"], g["c"]), g["iou"])
        assert torch.allclose(O.dice_per_class(g["z"], g["t"], g["c"]), g["dice"])


def test_state_dict_shapes_match_appendix_a():
    shapes = O.state_dict_shapes()
    assert len(shapes) == 182
    n_params = sum(int(np.prod(s)) for k, s in shapes.items()
                   if not k.endswith(("running_mean", "running_var", "num_batches_tracked")))
    assert n_params == 17_612_458
    assert len(O.state_dict_shapes(attention=False)) == 110
    assert len(O.state_dict_shapes(deep_supervision=True)) == 188
    assert len(O.state_dict_shapes(bilinear=False)) == 190


def test_io_cases():
    """prepare_slices / predict_mask against apply_basic_transforms, preprocess_image and
    postprocess_mask of the reference (fixture written by oracle/gen_golden.py:io_case)."""
    g = _load("io.pt")
    x, t = O.prepare_slices(g["images"].numpy(), g["labels"].numpy(), g["flags"].numpy())
    assert torch.equal(x, g["x"]) and torch.equal(t, g["t"])
    xp, none = O.prepare_slices(g["images"][:1].numpy(), requantize=False)
    assert none is None and torch.equal(xp, g["x_predict"])
    # the transform's second uint8 round trip changes no grey level
    lv = np.arange(256, dtype=np.uint8).reshape(1, 16, 16)
    assert torch.equal(O.prepare_slices(lv, requantize=True)[0], O.prepare_slices(lv, requantize=False)[0])
    m, c = O.predict_mask(g["z"], g["threshold"])
    assert np.array_equal(m, g["mask"].numpy()) and np.array_equal(c, g["positives"].numpy())
