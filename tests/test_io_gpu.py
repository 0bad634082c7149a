"""GPU: the two ends of the forward path (include/unetb200.h, "either side of the forward pass")
against the oracle and the fixture the reference's own functions produced (tests/golden/io.pt).
Integer / table work: bit-exact.  The thresholded mask is exact wherever the probability is not
within float rounding of the threshold."""
import os

import numpy as np
import pytest
import torch

from oracle import unet_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _slices(n, h, w, seed):
    rng = np.random.default_rng(seed)
    images = rng.integers(0, 256, (n, h, w), dtype=np.uint8)
    labels = rng.integers(0, 256, (n, h, w), dtype=np.uint8)
    return images, labels


def test_prepare_batch_golden():
    from unet.data import prepare_batch
    g = torch.load(os.path.join(GOLD, "io.pt"), weights_only=False)
    x, t = prepare_batch(g["images"].cuda(), g["labels"].cuda(), g["flags"].cuda())
    assert torch.equal(x.cpu(), g["x"]) and torch.equal(t.cpu(), g["t"])
    xp, none = prepare_batch(g["images"][:1].cuda())
    assert none is None and torch.equal(xp.cpu(), g["x_predict"])


@pytest.mark.parametrize("n,h,w", [(4, 512, 512), (3, 64, 48), (2, 33, 47), (1, 16, 16), (5, 7, 160)])
@pytest.mark.parametrize("mean,std", [(0.5, 0.5), (0.485, 0.229)])
def test_prepare_batch_vs_oracle(n, h, w, mean, std):
    from unet.data import prepare_batch
    images, labels = _slices(n, h, w, seed=n * 1000 + w)
    flags = (np.arange(n) % 4).astype(np.uint8)           # none, h, v, both
    ox, ot = O.prepare_slices(images, labels, flags, mean, std)
    x, t = prepare_batch(torch.from_numpy(images).cuda(), torch.from_numpy(labels).cuda(),
                         torch.from_numpy(flags).cuda(), mean, std)
    assert x.shape == (n, 1, h, w) and x.dtype == torch.float32 and t.dtype == torch.int64
    assert torch.equal(x.cpu(), ox) and torch.equal(t.cpu(), ot)
    # (N,1,H,W) input, no labels, no flags
    x2, t2 = prepare_batch(torch.from_numpy(images).cuda().unsqueeze(1), None, None, mean, std)
    assert t2 is None and torch.equal(x2.cpu(), O.prepare_slices(images, None, None, mean, std)[0])


def test_prepare_batch_properties_full_size():
    """512^2 batch 32: flipping twice is the identity; a flipped batch is the flip of the batch."""
    from unet.data import prepare_batch
    images, labels = _slices(32, 512, 512, seed=3)
    di, dl = torch.from_numpy(images).cuda(), torch.from_numpy(labels).cuda()
    x0, t0 = prepare_batch(di, dl)
    for bits, dims in ((1, (-1,)), (2, (-2,)), (3, (-2, -1))):
        f = torch.full((32,), bits, dtype=torch.uint8, device="cuda")
        x1, t1 = prepare_batch(di, dl, f)
        assert torch.equal(x1, x0.flip(dims)) and torch.equal(t1, t0.flip(dims))
    assert int(t0.sum()) == int((labels > 127).sum())
    assert float(x0.min()) >= -1.0 and float(x0.max()) <= 1.0


@pytest.mark.parametrize("n,h,w,thr", [(2, 48, 48, 0.6), (4, 512, 512, 0.5), (3, 31, 45, 0.3), (1, 8, 8, 0.5)])
def test_predict_mask_vs_oracle(n, h, w, thr):
    from unet import kernels as K
    g = torch.Generator().manual_seed(h * w + n)
    z = 3 * torch.randn(n, 2, h, w, generator=g)
    z[0, :, 0, :4] = 0.0          # exact ties: p = 0.5
    om, oc = O.predict_mask(z, thr)
    mask, pos = K.predict_mask(z.cuda(), thr)
    mask, pos = mask.cpu().numpy(), pos.cpu().numpy()
    assert set(np.unique(mask)) <= {0, 255}
    prob = torch.softmax(z.double(), 1)[:, 1].numpy()
    differ = mask != om
    assert not differ[np.abs(prob - thr) > 1e-6].any()
    assert differ.mean() <= 1e-5
    assert np.array_equal(pos, (mask > 127).reshape(n, -1).sum(1))
    assert np.abs(pos - oc).max() <= differ.reshape(n, -1).sum(1).max()


def test_predict_mask_golden():
    from unet import kernels as K
    g = torch.load(os.path.join(GOLD, "io.pt"), weights_only=False)
    mask, pos = K.predict_mask(g["z"].cuda(), g["threshold"])
    assert torch.equal(mask.cpu(), g["mask"]) and torch.equal(pos.cpu().long(), g["positives"])


def test_device_batch_pipeline():
    """Five uint8 batches through the double-buffered pipeline equal the oracle's host transform,
    with the flip bits the pipeline drew; float batches pass through unchanged."""
    from unet.data import DeviceBatchPipeline
    from unet.data.device import draw_flags
    batches = [tuple(torch.from_numpy(a) for a in _slices(4, 64, 64, seed=50 + i)) for i in range(5)]
    pipe = DeviceBatchPipeline(batches, "cuda", hflip_prob=0.5, vflip_prob=0.3, seed=9)
    gen = torch.Generator().manual_seed(9)
    assert len(pipe) == 5
    seen = 0
    for (x, t), (img, lab) in zip(pipe, batches):
        flags = draw_flags(4, gen, 0.5, 0.3)
        ox, ot = O.prepare_slices(img.numpy(), lab.numpy(), flags.numpy())
        # consume on the current stream, as a training step would
        assert torch.equal(x.cpu(), ox) and torch.equal(t.cpu(), ot)
        seen += 1
    assert seen == 5
    # a second epoch re-iterates the loader
    assert sum(1 for _ in pipe) == 5
    # the reference DataLoader's own output (fp32 images, int64 masks) is only copied
    fb = [(torch.randn(2, 1, 32, 32), torch.randint(0, 2, (2, 32, 32))) for _ in range(3)]
    for (x, t), (hx, ht) in zip(DeviceBatchPipeline(fb, "cuda"), fb):
        assert x.is_cuda and torch.equal(x.cpu(), hx) and torch.equal(t.cpu(), ht)


def test_pipeline_feeds_trainer_and_engine_predicts():
    """uint8 slices -> pipeline -> BatchShardedTrainer.step, then InferenceEngine.predict on raw
    slices equals thresholding the model's own logits."""
    from unet import kernels as K
    from unet.data import DeviceBatchPipeline
    from unet.inference import InferenceEngine
    from unet.models import AttentionUNet
    from unet.optim import FusedAdamW
    from unet.parallel import BatchShardedTrainer
    from unet.utils.loss import DiceBCELoss
    torch.manual_seed(0)
    model = AttentionUNet(1, 2, True, 32).cuda()
    trainer = BatchShardedTrainer(model, DiceBCELoss(), FusedAdamW(model.parameters(), lr=1e-3), grad_clip=1.0,
                                  cuda_graph=False)
    batches = [tuple(torch.from_numpy(a) for a in _slices(2, 64, 64, seed=70 + i)) for i in range(3)]
    losses = [float(trainer.step(x, t)) for x, t in DeviceBatchPipeline(batches, "cuda", hflip_prob=0.5)]
    assert all(np.isfinite(v) for v in losses)
    engine = InferenceEngine(model, cuda_graph=False)
    raw = batches[0][0].cuda()
    mask, ratio = engine.predict(raw, threshold=0.4)
    x, _ = K.prepare_batch(raw)
    with torch.no_grad():
        logits = model.eval()(x)
    ref_mask, ref_pos = K.predict_mask(logits.contiguous(), 0.4)
    assert torch.equal(mask, ref_mask)
    assert torch.allclose(ratio, ref_pos.double() / (64 * 64))
    om, _ = O.predict_mask(logits.float().cpu(), 0.4)
    assert (mask.cpu().numpy() != om).mean() <= 1e-4
