"""Per-module and end-to-end GPU parity of the drop-in modules against the oracle.

Protocol: the discriminating tests are *per module* — the CUDA module
and the fp32 oracle block get identical inputs, weights and upstream gradients.  bf16
storage / operands with fp32 accumulation give ~3e-3 relative error per module; the
tolerances below are the ones BASELINE.json states for bf16 mode:
  outputs / input gradients : relative L2 error <= 2e-2
  parameter gradients       : cosine similarity >= 0.999
  end-to-end eval logits    : relative L2 <= 2e-2, thresholded-mask agreement >= 99.9 %
"""
import copy

import pytest
import torch

from oracle import unet_oracle as O
from parity_log import record

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def cosine(a, b):
    a, b = a.detach().float().cpu().flatten(), b.detach().float().cpu().flatten()
    return (torch.dot(a, b) / (a.norm() * b.norm() + 1e-20)).item()


def randomise(module, seed):
    """Non-trivial parameters and BN buffers, deterministic per key."""
    g = torch.Generator().manual_seed(seed)
    sd = module.state_dict()
    for k, v in sd.items():
        if k.endswith("num_batches_tracked"):
            continue
        if k.endswith("running_var"):
            v.copy_(0.5 + torch.rand(v.shape, generator=g))
        elif k.endswith("running_mean"):
            v.copy_(0.1 * torch.randn(v.shape, generator=g))
        elif v.dim() == 4:
            fan = v.shape[1] * v.shape[2] * v.shape[3]
            v.copy_((torch.rand(v.shape, generator=g) * 2 - 1) / fan ** 0.5)
        elif k.endswith("weight"):
            v.copy_(0.75 + 0.5 * torch.rand(v.shape, generator=g))
        else:
            v.copy_(0.1 * torch.randn(v.shape, generator=g))
    module.load_state_dict(sd)
    return module


def oracle_sd(module, prefix="m"):
    return {f"{prefix}.{k}": v.detach().clone().cpu() for k, v in module.state_dict().items()}


def run_oracle(fn, sd, inputs, grad_out, training):
    """Run an oracle block with autograd; returns (out, input grads, param grads, updated sd)."""
    sd = {k: v.clone() for k, v in sd.items()}
    params = {k: v.requires_grad_(True) for k, v in sd.items()
              if v.is_floating_point() and not k.endswith(("running_mean", "running_var"))}
    ins = [x.clone().requires_grad_(True) for x in inputs]
    out = fn(*ins, sd, training)
    out.backward(grad_out)
    return out.detach(), [x.grad for x in ins], {k: p.grad for k, p in params.items()}, sd


def check_module(module, oracle_fn, inputs, training, seed=0, input_needs_grad=True, out_tol=2e-2, tag=None):
    """``tag``: name under which every measured number goes to the parity log (default: the test's)."""
    import os
    tag = tag or os.environ.get("PYTEST_CURRENT_TEST", "module").split("::")[-1].split(" ")[0]
    log = {}
    torch.manual_seed(seed)
    module = randomise(module, seed + 1)
    module.train(training)
    sd = oracle_sd(module)
    # bf16-representable inputs so both sides see identical values
    inputs = [x.to(torch.bfloat16).float() for x in inputs]
    cuda_mod = copy.deepcopy(module).cuda()
    cins = [x.cuda().requires_grad_(input_needs_grad) for x in inputs]
    out = cuda_mod(*cins)
    g = torch.Generator().manual_seed(seed + 2)
    grad_out = torch.randn(out.shape, generator=g).to(torch.bfloat16).float()
    out.backward(grad_out.cuda().to(out.dtype))
    ref_out, ref_in, ref_par, ref_sd = run_oracle(oracle_fn, sd, inputs, grad_out, training)

    assert out.shape == ref_out.shape
    # Second oracle run with the product's rounding points (bf16 conv operands, bf16 stored
    # activations): ReLU / max-pool decisions are then taken on identical values, which makes
    # this the tight, bug-finding comparison.  Against the pure-fp32 oracle the error is
    # dominated by sign flips of near-zero pre-activations (a flip fraction f costs sqrt(f)
    # relative L2): the two oracle runs differ from each other by the same 4-7 % on input
    # gradients / 0.997-0.9999 parameter-gradient cosine under white-noise upstream gradients,
    # so the fp32 gates are: output rel-L2 <= 2e-2 (BASELINE.json) and cosine >= 0.995.
    with O.bf16_storage():
        st_out, st_in, st_par, _ = run_oracle(oracle_fn, sd, inputs, grad_out, training)
    problems = []
    report = [f"out rel-L2: fp32 {rel_l2(out, ref_out):.3e}  bf16-model {rel_l2(out, st_out):.3e}"]
    log.update(out_rel_l2_vs_fp32=rel_l2(out, ref_out), out_rel_l2_vs_bf16_model=rel_l2(out, st_out),
               out_rel_l2_floor_bf16_model_vs_fp32=rel_l2(st_out, ref_out))
    if rel_l2(out, ref_out) > out_tol:
        problems.append(f"output rel-L2 vs fp32 {rel_l2(out, ref_out):.3e}")
    if rel_l2(out, st_out) > 5e-3:
        problems.append(f"output rel-L2 vs bf16-model {rel_l2(out, st_out):.3e}")
    if input_needs_grad:
        for i, (ci, ri, si) in enumerate(zip(cins, ref_in, st_in)):
            c, e = cosine(ci.grad, ri), rel_l2(ci.grad, si)
            report.append(f"in{i}: cos(fp32) {c:.5f} rel-L2(bf16-model) {e:.3e}")
            log[f"in{i}_grad"] = {"cos_vs_fp32": c, "rel_l2_vs_bf16_model": e, "cos_floor_bf16_model_vs_fp32": cosine(si, ri)}
            if c < 0.995:
                problems.append(f"input {i} grad cosine vs fp32 oracle {c:.5f}")
            if e > 2e-2:
                problems.append(f"input {i} grad rel-L2 vs bf16-model oracle {e:.3e}")
    # A BatchNorm2d(1) (the gate's psi.1) has one-element weight / bias gradients: a cosine is +-1 and the
    # bias gradient is a plain sum of signed per-pixel terms that nearly cancels under white-noise upstream
    # gradients, so its relative error is unbounded.  Such pairs are compared as ONE vector (dgamma, dbeta).
    named = dict(cuda_mod.named_parameters())
    groups = []
    for name, p in named.items():
        if p.numel() == 1 and name.endswith(".weight") and name[:-6] + "bias" in named and named[name[:-6] + "bias"].numel() == 1:
            groups.append((name[:-7] + ".{weight,bias}", [name, name[:-6] + "bias"]))
        elif p.numel() == 1 and name.endswith(".bias") and name[:-4] + "weight" in named and named[name[:-4] + "weight"].numel() == 1:
            continue
        else:
            groups.append((name, [name]))
    for name, members in groups:
        if any(ref_par["m." + m] is None for m in members):
            continue
        got = torch.cat([named[m].grad.detach().float().cpu().flatten() for m in members])
        r = torch.cat([ref_par["m." + m].flatten() for m in members])
        r2 = torch.cat([st_par["m." + m].flatten() for m in members])
        if r.norm() == 0:
            continue
        c, c2 = cosine(got, r), cosine(got, r2)
        scale = (got.norm() / r2.norm()).item()
        report.append(f"{name}: cos fp32 {c:.5f} bf16-model {c2:.5f} norm ratio {scale:.4f}")
        log[f"grad {name}"] = {"cos_vs_fp32": c, "cos_vs_bf16_model": c2, "cos_floor_bf16_model_vs_fp32": cosine(r2, r),
                               "norm_ratio": scale}
        if len(members) > 1:
            log[f"grad {name}"]["values"] = {"got": got.tolist(), "bf16_model": r2.tolist(), "fp32": r.tolist()}
        # Gates: cosine >= 0.995 vs fp32 and >= 0.999 vs the bf16 model (north_star's figure) — unless the two
        # ORACLES already disagree more than that on this tensor (a sum that nearly cancels: the gate's psi
        # BatchNorm under white-noise upstream gradients at 512 channels).  Then the product has to sit much
        # closer to the bf16 model than the bf16 model sits to fp32, and both numbers are logged.
        floor = cosine(r2, r)
        near_model = (got - r2).norm().item() <= 0.25 * (r2 - r).norm().item()
        log[f"grad {name}"]["dist_to_bf16_model_over_model_to_fp32"] = ((got - r2).norm() / ((r2 - r).norm() + 1e-30)).item()
        if c < 0.995 and c < floor - 1e-3 and not near_model:
            problems.append(f"{name}: grad cosine vs fp32 {c:.5f} (oracle-vs-oracle floor {floor:.5f})")
        if c2 < 0.999 and not near_model:
            problems.append(f"{name}: grad cosine vs bf16-model {c2:.5f}")
        if abs(scale - 1) > 2e-2 and not near_model:
            problems.append(f"{name}: grad norm ratio {scale:.4f}")
    if training:
        for k, v in cuda_mod.state_dict().items():
            if k.endswith(("running_mean", "running_var")):
                if not torch.allclose(v.cpu(), ref_sd["m." + k], rtol=2e-2, atol=2e-3):
                    problems.append(f"{k} mismatch")
            if k.endswith("num_batches_tracked") and int(v) != int(ref_sd["m." + k]):
                problems.append(f"{k}: {int(v)} vs {int(ref_sd['m.' + k])}")
    print("\n".join(report))
    record(tag, shapes=[list(x.shape) for x in inputs], training=training, problems=problems, **log)
    assert not problems, "; ".join(problems)


def _x(shape, seed):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed))


@pytest.mark.parametrize("training", [True, False])
def test_double_conv(training):
    from unet.models.layers import DoubleConv
    check_module(DoubleConv(64, 128), lambda x, sd, tr: O.double_conv(x, sd, "m", tr),
                 [_x((2, 64, 24, 20), 1)], training)


def test_double_conv_stem():
    from unet.models.layers import DoubleConv
    check_module(DoubleConv(1, 64), lambda x, sd, tr: O.double_conv(x, sd, "m", tr),
                 [_x((2, 1, 32, 32), 2)], True, input_needs_grad=False)


@pytest.mark.parametrize("training", [True, False])
def test_down(training):
    from unet.models.layers import Down
    check_module(Down(64, 128), lambda x, sd, tr: O.down(x, sd, "m", tr), [_x((2, 64, 32, 32), 3)], training)


def test_down_odd_size():
    from unet.models.layers import Down
    check_module(Down(32, 64), lambda x, sd, tr: O.down(x, sd, "m", tr), [_x((1, 32, 19, 27), 4)], True)


@pytest.mark.parametrize("bilinear", [True, False])
def test_up(bilinear):
    from unet.models.layers import Up
    c_low = 64 if bilinear else 128
    check_module(Up(128, 64, bilinear),
                 lambda x1, x2, sd, tr: O.up_block(x1, x2, sd, "m", False, bilinear, tr),
                 [_x((2, c_low, 8, 8), 5), _x((2, 64, 16, 16), 6)], True)


def test_up_padded():
    from unet.models.layers import Up
    check_module(Up(128, 64, True), lambda x1, x2, sd, tr: O.up_block(x1, x2, sd, "m", False, True, tr),
                 [_x((1, 64, 6, 9), 7), _x((1, 64, 13, 19), 8)], True)


@pytest.mark.parametrize("training", [True, False])
def test_attention_gate(training):
    from unet.models.layers import AttentionGate
    check_module(AttentionGate(128, 128), lambda g, x, sd, tr: O.attention_gate(g, x, sd, "m", tr),
                 [_x((2, 128, 8, 8), 9), _x((2, 128, 16, 16), 10)], training)


def test_attention_gate_small_inter():
    from unet.models.layers import AttentionGate
    check_module(AttentionGate(64, 64), lambda g, x, sd, tr: O.attention_gate(g, x, sd, "m", tr),
                 [_x((1, 64, 16, 16), 11), _x((1, 64, 32, 32), 12)], True)


def test_attention_up():
    from unet.models.layers import AttentionUp
    check_module(AttentionUp(256, 64, True),
                 lambda x1, x2, sd, tr: O.up_block(x1, x2, sd, "m", True, True, tr),
                 [_x((2, 128, 8, 8), 13), _x((2, 128, 16, 16), 14)], True)


def test_out_conv():
    from unet.models.layers import OutConv
    import torch.nn.functional as F
    check_module(OutConv(64, 2), lambda x, sd, tr: F.conv2d(x, sd["m.conv.weight"], sd["m.conv.bias"]),
                 [_x((2, 64, 16, 16), 15)], True, out_tol=1e-3)


# --------------------------------------------------------------------------- whole network
def _build(attention, bf, seed, **kw):
    from unet.models import AttentionUNet, UNet
    cfg = dict(n_channels=1, n_classes=2, bilinear=True, base_features=bf, attention=attention)
    cfg.update(kw)
    sd = O.synthetic_state_dict(seed, **cfg)
    mk = dict(n_channels=1, n_classes=2, bilinear=cfg["bilinear"], base_features=bf)
    if attention and cfg.get("deep_supervision"):
        mk["deep_supervision"] = True
    model = (AttentionUNet if attention else UNet)(**mk)
    model.load_state_dict(sd, strict=True)
    return model, sd, cfg


@pytest.mark.parametrize("attention,bf", [(True, 64), (False, 32), (True, 32)])
def test_eval_end_to_end(attention, bf):
    model, sd, cfg = _build(attention, bf, 21)
    x, t = O.synthetic_batch(2, 64, 64, seed=5)
    model = model.cuda().eval()
    with torch.no_grad():
        logits = model(x.cuda())
    ocfg = {k: v for k, v in cfg.items() if k in ("bilinear", "deep_supervision")}
    ref = O.unet_forward(x, sd, attention=attention, training=False, **ocfg)
    assert logits.dtype == torch.float32 and logits.shape == ref.shape
    m_got = torch.softmax(logits.cpu(), 1)[:, 1] > 0.5
    m_ref = torch.softmax(ref, 1)[:, 1] > 0.5
    record(f"test_eval_end_to_end[{attention}-{bf}]", logits_rel_l2_vs_fp32=rel_l2(logits, ref),
           mask_agreement=(m_got == m_ref).float().mean().item(), shape=list(x.shape))
    assert rel_l2(logits, ref) <= 2e-2, f"eval logits rel-L2 {rel_l2(logits, ref):.3e}"
    assert (m_got == m_ref).float().mean().item() >= 0.999


def test_train_step_end_to_end_report():
    """bf16 train-mode end-to-end at random init is chaotic (PyTorch's own
    bf16 autocast differs from fp32 by 13 % logits / median grad cosine 0.94): report the
    numbers, gate only on sanity."""
    from unet.utils.loss import DiceBCELoss
    model, sd, cfg = _build(True, 32, 22)
    x, t = O.synthetic_batch(2, 64, 64, seed=6, fg_fraction=0.05)
    model = model.cuda().train()
    logits = model(x.cuda())
    loss = DiceBCELoss()(logits, t.cuda())
    loss.backward()
    ref_loss, ref_logits, ref_grads = O.train_grads(x, t, O.clone_state(sd), attention=True)
    e = rel_l2(logits, ref_logits)
    cos = sorted(cosine(p.grad, ref_grads[k]) for k, p in model.named_parameters())
    print(f"train e2e: logits rel-L2 {e:.3e}, loss {loss.item():.5f} vs {ref_loss.item():.5f}, "
          f"grad cosine min {cos[0]:.4f} median {cos[len(cos) // 2]:.4f}")
    record("test_train_step_end_to_end_report", logits_rel_l2_vs_fp32=e, loss=loss.item(), loss_oracle=ref_loss.item(),
           grad_cos_min=cos[0], grad_cos_median=cos[len(cos) // 2],
           note="bf16 train mode at random init is chaotic (SURVEY App. C: torch's own bf16 autocast vs fp32 gives "
                "1.3e-1 / median cosine 0.94); reported, gated on sanity only")
    assert torch.isfinite(loss).item() and e < 0.5 and cos[len(cos) // 2] > 0.8
    assert abs(loss.item() - ref_loss.item()) < 0.1


def test_state_dict_roundtrip_and_deepcopy():
    model, sd, _ = _build(True, 32, 23, deep_supervision=True)
    got = model.state_dict()
    assert list(got.keys()) == list(sd.keys())
    for k in sd:
        assert got[k].shape == sd[k].shape and got[k].dtype == sd[k].dtype, k
    clone = copy.deepcopy(model).cuda()
    clone.load_state_dict(model.state_dict(), strict=True)


def test_cpu_tensor_raises():
    from unet.models import UNet
    with pytest.raises(RuntimeError, match="CUDA"):
        UNet(base_features=16)(torch.zeros(1, 1, 32, 32))


def test_cuda_graph_step_equals_eager():
    """The captured-and-replayed training step is the same arithmetic as the eager one."""
    from unet.models import AttentionUNet
    from unet.parallel import BatchShardedTrainer
    from unet.utils.loss import DiceBCELoss
    x, t = O.synthetic_batch(2, 64, 64, seed=9, fg_fraction=0.05)
    x, t = x.cuda(), t.cuda()
    results = []
    for use_graph in (False, True):
        torch.manual_seed(3)
        model = AttentionUNet(1, 2, True, 32).cuda()
        opt = torch.optim.AdamW(model.parameters(), lr=1e-3, capturable=True, foreach=True)
        tr = BatchShardedTrainer(model, DiceBCELoss(), opt, grad_clip=1.0, cuda_graph=use_graph, graph_warmup=2)
        losses = [tr.step(x, t).item() for _ in range(6)]
        results.append((losses, [p.detach().clone() for p in model.parameters()],
                        int(model.inc.double_conv[1].num_batches_tracked)))
    (l0, p0, n0), (l1, p1, n1) = results
    assert n0 == n1 == 6
    assert all(abs(a - b) <= 1e-6 * abs(a) for a, b in zip(l0, l1)), (l0, l1)
    for a, b in zip(p0, p1):
        assert torch.equal(a, b)


@pytest.mark.parametrize("use_graph", [False, True])
def test_weight_gradients_on_the_side_stream_change_nothing(use_graph, monkeypatch):
    """ops.SideStream: the weight-gradient GEMMs forked next to the BatchNorm-backward passes produce the same
    bits as the one-stream order (same kernels, same split-K; only the schedule differs), eager and replayed,
    over enough steps that a premature fold or a recycled dy would show."""
    from unet import ops
    from unet.models import AttentionUNet
    from unet.parallel import BatchShardedTrainer
    from unet.optim import FusedAdamW
    from unet.utils.loss import DiceBCELoss
    x, t = O.synthetic_batch(2, 96, 64, seed=11, fg_fraction=0.05)
    x, t = x.cuda(), t.cuda()
    results = []
    for side in ("0", "1"):
        monkeypatch.setenv("UB2_WGRAD_SIDE", side)
        torch.manual_seed(5)
        model = AttentionUNet(1, 2, True, 32).cuda()
        tr = BatchShardedTrainer(model, DiceBCELoss(), FusedAdamW(model.parameters(), lr=1e-3), grad_clip=1.0,
                                 cuda_graph=use_graph, graph_warmup=2)
        assert (tr._wgrad_side(x) is not None) == (side == "1")
        losses = [tr.step(x, t).item() for _ in range(8)]
        assert ops.WGRAD_SIDE is None                       # installed for the duration of backward only
        results.append((losses, [p.detach().clone() for p in model.parameters()]))
    (l0, p0), (l1, p1) = results
    assert l0 == l1, (l0, l1)
    for a, b in zip(p0, p1):
        assert torch.equal(a, b)


def test_accumulation_steps_graph_equals_eager_and_autograd():
    """accumulation_steps=2 (train.py:127-147 on one GPU): replayed graphs (one per micro-batch phase)
    == eager, and the accumulated gradient equals plain autograd's sum over the two micro-batches."""
    from unet.models import AttentionUNet
    from unet.optim import FusedAdamW
    from unet.parallel import BatchShardedTrainer
    from unet.utils.loss import DiceBCELoss
    data = [tuple(v.cuda() for v in O.synthetic_batch(2, 64, 64, seed=20 + i, fg_fraction=0.05)) for i in range(2)]
    results = []
    for use_graph in (False, True):
        torch.manual_seed(3)
        model = AttentionUNet(1, 2, True, 32).cuda()
        start = [p.detach().clone() for p in model.parameters()]
        opt = FusedAdamW(model.parameters(), lr=1e-3)
        tr = BatchShardedTrainer(model, DiceBCELoss(), opt, grad_clip=1.0, cuda_graph=use_graph, graph_warmup=1,
                                 accumulation_steps=2)
        losses = []
        for step in range(4):
            for m in range(2):
                losses.append(tr.step(*data[m]).item())
                if step == 0 and m == 0:   # nothing moves before the second micro-batch
                    assert all(torch.equal(a, b) for a, b in zip(start, model.parameters()))
        results.append((losses, [p.detach().clone() for p in model.parameters()],
                        int(model.inc.double_conv[1].num_batches_tracked)))
    (l0, p0, n0), (l1, p1, n1) = results
    assert n0 == n1 == 8
    assert all(abs(a - b) <= 1e-6 * abs(a) for a, b in zip(l0, l1)), (l0, l1)
    for a, b in zip(p0, p1):
        assert torch.equal(a, b)
    # gradient of one accumulated step against autograd's own accumulation (no sink, no buckets)
    torch.manual_seed(3)
    ref = AttentionUNet(1, 2, True, 32).cuda()
    crit = DiceBCELoss()
    ref.train()
    for m in range(2):
        (crit(ref(data[m][0]), data[m][1]) / 2).backward()
    torch.manual_seed(3)
    model = AttentionUNet(1, 2, True, 32).cuda()
    tr = BatchShardedTrainer(model, crit, FusedAdamW(model.parameters(), lr=0.0, weight_decay=0.0), accumulation_steps=2)
    for m in range(2):
        tr.step(*data[m])
    for (k, a), b in zip(model.named_parameters(), ref.parameters()):
        assert torch.allclose(a.grad, b.grad, rtol=1e-4, atol=1e-6 * float(b.grad.abs().max()) + 1e-9), k


def test_direct_gradient_sink_equals_autograd():
    """With a gradient sink installed the kernels accumulate straight into param.grad and autograd
    sees None: the result is bit-identical to the AccumulateGrad path, every parameter reported once."""
    from unet import ops
    from unet.models import AttentionUNet
    from unet.utils.loss import DiceBCELoss
    x, t = O.synthetic_batch(2, 64, 64, seed=11, fg_fraction=0.05)
    x, t = x.cuda(), t.cuda()
    torch.manual_seed(5)
    model = AttentionUNet(1, 2, True, 32).cuda().train()
    crit = DiceBCELoss()
    crit(model(x), t).backward()
    ref = [p.grad.clone() for p in model.parameters()]
    for p in model.parameters():
        p.grad = torch.zeros_like(p)
    # BatchNorm buffers moved in the first pass; batch statistics do not depend on them
    seen = []
    ops.GRAD_SINK = seen.append
    try:
        crit(model(x), t).backward()
    finally:
        ops.GRAD_SINK = None
    params = list(model.parameters())
    assert sorted(id(p) for p in seen) == sorted(id(p) for p in params)
    for p, r in zip(params, ref):
        assert torch.equal(p.grad, r)
    # accumulation semantics: a second backward adds
    ops.GRAD_SINK = lambda p: None
    try:
        crit(model(x), t).backward()
    finally:
        ops.GRAD_SINK = None
    for p, r in zip(params, ref):
        assert torch.allclose(p.grad, 2 * r, rtol=1e-5, atol=1e-6)


def test_inference_engine_graph_equals_eager():
    """InferenceEngine (weight packs built once, forward replayed as a CUDA graph) == model.eval()(x)."""
    from unet.inference import InferenceEngine
    from unet.models import AttentionUNet
    torch.manual_seed(4)
    model = AttentionUNet(1, 2, True, 32).cuda().eval()
    engine = InferenceEngine(model)
    for n in (1, 2):
        x, _ = O.synthetic_batch(n, 64, 64, seed=20 + n)
        with torch.no_grad():
            ref = model(x.cuda())
        outs = [engine(x.pin_memory()).clone() for _ in range(4)]   # eager warm-ups, capture, replays
        for o in outs:
            assert torch.equal(o, ref)
    with torch.no_grad():   # new weights: refresh() rebuilds the packs and drops the graphs
        for p in model.parameters():
            p.mul_(1.01)
    engine.refresh()
    x, _ = O.synthetic_batch(1, 64, 64, seed=30)
    with torch.no_grad():
        ref = model(x.cuda())
    assert torch.equal(engine(x).clone(), ref)


def test_deep_supervision_outputs_and_loss():
    """deep_supervision=True in training mode returns [main, ds1, ds2, ds3] at the input size
    (unet.py:204-211); DeepSupervisionLoss (loss.py:194-229) of them matches the oracle's; in eval
    mode a single tensor comes back."""
    from unet.utils.loss import DeepSupervisionLoss, DiceBCELoss
    model, sd, cfg = _build(True, 32, 31, deep_supervision=True)
    x, t = O.synthetic_batch(2, 64, 64, seed=7, fg_fraction=0.05)
    model = model.cuda().train()
    outs = model(x.cuda())
    assert isinstance(outs, list) and len(outs) == 4 and all(o.shape == (2, 2, 64, 64) for o in outs)
    loss = DeepSupervisionLoss(DiceBCELoss())(outs, t.cuda())
    ref = O.deep_supervision_loss([o.detach().cpu().float() for o in outs], t)
    assert abs(loss.item() - ref.item()) <= 1e-4 * abs(ref.item()) + 1e-6
    loss.backward()
    heads = [p for n, p in model.named_parameters() if n.startswith("ds_out")]
    assert len(heads) == 6 and all(p.grad is not None and torch.isfinite(p.grad).all() and p.grad.abs().sum() > 0
                                   for p in heads)
    ref_outs = O.unet_forward(x, O.clone_state(sd), attention=True, deep_supervision=True, training=True)
    for o, r in zip(outs, ref_outs):   # train mode at random init: loose gate
        assert rel_l2(o, r) < 0.5
    model.eval()
    with torch.no_grad():
        assert torch.is_tensor(model(x.cuda()))


def test_eval_end_to_end_transposed_conv_variant():
    """bilinear=False (ConvTranspose2d up-sampling, layers.py:81, :217-221), eval mode."""
    model, sd, cfg = _build(True, 32, 33, bilinear=False)
    x, _ = O.synthetic_batch(2, 64, 64, seed=8)
    model = model.cuda().eval()
    with torch.no_grad():
        logits = model(x.cuda())
    ref = O.unet_forward(x, sd, attention=True, bilinear=False, training=False)
    assert rel_l2(logits, ref) <= 2e-2, f"eval logits rel-L2 {rel_l2(logits, ref):.3e}"


def test_full_size_step_is_deterministic_and_finite():
    """BASELINE configs[1] size (4 x 1 x 512 x 512, base 64): two identical training steps from the
    same state give bit-identical losses and gradients (fixed-order reductions, no atomics), and the
    batch-statistics identity holds: the raw conv outputs feeding every BatchNorm are normalised to
    zero mean / unit variance, so the BN running means move by exactly momentum * batch mean."""
    from unet.models import AttentionUNet
    from unet.utils.loss import DiceBCELoss
    x, t = O.synthetic_batch(4, 512, 512, seed=1234)
    x, t = x.cuda(), t.cuda()
    runs = []
    for _ in range(2):
        torch.manual_seed(42)
        model = AttentionUNet(1, 2, True, 64).cuda().train()
        loss = DiceBCELoss()(model(x), t)
        loss.backward()
        runs.append((loss.item(), [p.grad.clone() for p in model.parameters()],
                     model.inc.double_conv[1].running_mean.clone()))
    (l0, g0, rm0), (l1, g1, rm1) = runs
    assert l0 == l1 and torch.isfinite(torch.tensor(l0))
    for a, b in zip(g0, g1):
        assert torch.equal(a, b) and torch.isfinite(a).all()
    assert torch.equal(rm0, rm1) and rm0.abs().sum() > 0


@pytest.mark.parametrize("n,h,w,ci", [(2, 16, 16, 64), (1, 13, 19, 32), (2, 64, 64, 128), (1, 512, 512, 32)])
def test_gate_fused_eval_equals_two_passes(n, h, w, ci):
    """Inference runs psi + sigmoid + gating as ONE kernel (north_star (3)): bit-identical to gate_psi followed by
    gate_apply, which the training path and its parity tests use."""
    from unet import kernels as K
    g = torch.Generator(device="cuda").manual_seed(ci + h)
    rnd = lambda *s: torch.randn(*s, device="cuda", generator=g)
    hin, win = (h + 1) // 2, (w + 1) // 2
    q, xp, x = rnd(n, hin, win, ci).bfloat16(), rnd(n, h, w, ci).bfloat16(), rnd(n, h, w, 2 * ci).bfloat16()
    sg, sx = rnd(ci).abs() + 0.5, rnd(ci).abs() + 0.5
    hg, hx, wpsi = rnd(ci) * 0.3, rnd(ci) * 0.3, rnd(ci) * 0.2
    sp, hp = torch.tensor([0.7], device="cuda"), torch.tensor([-0.1], device="cuda")
    psi, _ = K.gate_psi(q, xp, sg, hg, sx, hx, wpsi, stats=False)
    ref, _ = K.gate_apply(psi, sp, hp, x, save_a=False)
    got = K.gate_fused_eval(q, xp, x, sg, hg, sx, hx, wpsi, sp, hp)
    assert torch.equal(got, ref)


@pytest.mark.parametrize("precision", ["bf16", "tf32"])
def test_base_features_16_gate_with_8_inter_channels(precision):
    """AttentionUNet(base_features=16): up4's gate has 8 inter-channels (unet/models/unet.py:137-164 accepts any
    base_features), below the tensor cores' multiple-of-16 granularity — the gate runs through a zero-padded view
    (models/layers.py:_PaddedGate).  Eval logits vs the oracle, one train step (loss, finite gradients of the right
    shapes, the gate's BatchNorm running statistics written back), state_dict layout unchanged."""
    import unet
    from unet.utils.loss import DiceBCELoss
    unet.set_precision(precision)
    try:
        model, sd, cfg = _build(True, 16, 41)
        assert model.up4.attention.W_g[0].weight.shape[0] == 8
        x, t = O.synthetic_batch(2, 64, 64, seed=12, fg_fraction=0.05)
        model = model.cuda().eval()
        with torch.no_grad():
            logits = model(x.cuda())
        ref = O.unet_forward(x, sd, attention=True, training=False)
        tol = 2e-2 if precision == "bf16" else 1e-3
        e = rel_l2(logits, ref)
        assert e <= tol, f"eval logits rel-L2 {e:.3e}"
        model.train()
        out = model(x.cuda())
        loss = DiceBCELoss()(out, t.cuda())
        loss.backward()
        ref_state = O.clone_state(sd)
        ref_loss, ref_logits, ref_grads = O.train_grads(x, t, ref_state, attention=True)
        cos = {k: cosine(p.grad, ref_grads[k]) for k, p in model.named_parameters() if p.numel() > 1}
        record(f"base_features=16 [{precision}]", eval_logits_rel_l2=e, loss=loss.item(), loss_oracle=ref_loss.item(),
               train_logits_rel_l2=rel_l2(out, ref_logits), grad_cos_min=min(cos.values()),
               grad_cos_gate8={k: v for k, v in cos.items() if k.startswith("up4.attention")})
        assert abs(loss.item() - ref_loss.item()) <= (5e-2 if precision == "bf16" else 1e-4) * abs(ref_loss.item())
        for k, p in model.named_parameters():
            assert p.grad is not None and p.grad.shape == p.shape and torch.isfinite(p.grad).all(), k
        gate_cos = [v for k, v in cos.items() if k.startswith("up4.attention")]
        # bf16 train mode end to end at random init is chaotic for EVERY parameter (test_train_step_end_to_end_report);
        # the fp32 mode is the discriminating check of the padded gate's arithmetic
        assert min(gate_cos) >= (0.8 if precision == "bf16" else 0.999), cos
        got_sd = model.state_dict()
        assert list(got_sd.keys()) == list(sd.keys())
        for k in ("up4.attention.W_g.1.running_mean", "up4.attention.W_x.1.running_var"):
            assert got_sd[k].shape == sd[k].shape
            assert torch.allclose(got_sd[k].cpu(), ref_state[k], rtol=5e-2, atol=5e-3), k
        assert int(got_sd["up4.attention.W_g.1.num_batches_tracked"]) == int(sd["up4.attention.W_g.1.num_batches_tracked"]) + 1
    finally:
        unet.set_precision("bf16")


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_model_on_second_device_while_first_is_current():
    """The reference's get_device('cuda:1') pattern: the model and its inputs live on cuda:1 while the process's
    current device stays cuda:0.  Every C-ABI call must run with the tensors' device current and on that device's
    stream (unet/_C.py), and the launchers' one-time state is per device ordinal (csrc/conv.h: PerDevice)."""
    from unet.models import AttentionUNet
    from unet.utils.loss import DiceBCELoss
    assert torch.cuda.current_device() == 0
    torch.manual_seed(5)
    m0 = AttentionUNet(1, 2, True, 32)
    m1 = copy.deepcopy(m0).to("cuda:1").train()
    m0 = m0.to("cuda:0").train()
    x, t = O.synthetic_batch(2, 64, 64, seed=3, fg_fraction=0.05)
    outs = []
    for m, dev in ((m0, "cuda:0"), (m1, "cuda:1"), (m0, "cuda:0")):
        if len(outs) == 2:
            for p in m.parameters():
                p.grad = None
        out = m(x.to(dev))
        assert out.device == torch.device(dev)
        loss = DiceBCELoss()(out, t.to(dev))
        loss.backward()
        torch.cuda.synchronize(dev)
        outs.append((out.detach().cpu(), [p.grad.detach().cpu() for p in m.parameters()]))
    assert torch.cuda.current_device() == 0
    (o0, g0), (o1, g1), _ = outs
    assert torch.equal(o0, o1)
    for a, b in zip(g0, g1):
        assert torch.equal(a, b)
