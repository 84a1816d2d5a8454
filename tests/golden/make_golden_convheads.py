"""Golden fixtures for the Conv1d heads from the LIVE reference (run in the authoring container):

    python tests/golden/make_golden_convheads.py

RestorerConv1d (models.py:661-716) and ClassifierConv1d (:865-902) in train mode and in eval mode.  The dropout keep-masks
torch draws inside the reference modules are CAPTURED with forward hooks on its nn.Dropout layers (output != 0 would miss
kept zeros, so the mask is recovered from output vs input) and stored, so that the oracle and the CUDA path can replay the
exact same masks.  Stored per case: inputs, masks, outputs, the gradients of every parameter and of the input for the
upstream gradient d_out, the BatchNorm buffers after the step.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import models as ref                      # noqa: E402
from oracle import iins_oracle as orc      # noqa: E402
from tests.golden.make_golden_common import conv_head_case_inputs   # noqa: E402


def capture_masks(module):
    masks, hooks = [], []

    def hook(mod, inp, out):
        x = inp[0]
        keep = torch.where(x != 0, (out != 0).float(), torch.ones_like(x))     # an exact zero input tells nothing: count as kept
        masks.append(keep.detach().clone())

    for m in module.modules():
        if isinstance(m, torch.nn.Dropout):
            hooks.append(m.register_forward_hook(hook))
    return masks, hooks


def run_case(store, prefix, kind, seed, batch, training, cfg):
    x, p, gen = conv_head_case_inputs(kind, seed, batch, cfg)
    if kind == "res":
        mod = ref.Restorer(code_shape=(cfg.range_dim, cfg.code_len), soft=False, filters=cfg.dim, conv_type=1, expand=False, net_type="Conv1d")
    else:
        mod = ref.Classifier(env_dim=cfg.env_dim, num_classes=cfg.num_classes, filters=16, net_type="Conv1d")
    assert list(mod.state_dict().keys()) == list(p.keys()), (list(mod.state_dict().keys()), list(p.keys()))
    mod.load_state_dict(p)
    mod.train(training)
    masks, hooks = capture_masks(mod)
    x = x.clone().requires_grad_(True)
    torch.manual_seed(seed + 5)
    out = mod(x)
    d_out = torch.randn(out.shape, generator=gen) * 0.1
    (out * d_out).sum().backward()
    for h in hooks:
        h.remove()
    store[prefix + "meta"] = np.array([seed, batch, int(training)], dtype=np.int64)
    store[prefix + "x"] = x.detach().numpy()
    store[prefix + "d_out"] = d_out.numpy()
    store[prefix + "out"] = out.detach().numpy()
    store[prefix + "d_x"] = x.grad.numpy()
    for i, m in enumerate(masks):
        store[prefix + f"mask{i}"] = m.numpy()
    for k, v in mod.named_parameters():
        store[prefix + "grad." + k] = np.zeros(0, dtype=np.float32) if v.grad is None else v.grad.numpy()
    for k, v in mod.state_dict().items():
        if "running" in k or "num_batches" in k:
            store[prefix + "buf." + k] = v.numpy()
    # the oracle must reproduce the reference on the same masks
    om = tuple(masks) if training else None
    fn = orc.restorer_conv1d if kind == "res" else orc.classifier_conv1d
    o, (rm, rv) = fn(p, x.detach(), om, training)
    assert torch.allclose(o, out.detach(), rtol=1e-5, atol=1e-6), float((o - out).abs().max())
    if training:
        key = [k for k in p if k.endswith("running_mean")][0]
        assert torch.allclose(rm, mod.state_dict()[key], rtol=1e-5, atol=1e-6)
        assert torch.allclose(rv, mod.state_dict()[key.replace("mean", "var")], rtol=1e-5, atol=1e-6)


def run_soft_case(store, prefix, seed, batch, cfg):
    """RestorerLinear with soft=True (models.py:634-655): the (B, 1) x (B,) broadcast makes the output (B, B)."""
    gen = torch.Generator().manual_seed(seed)
    p = orc.init_params(orc.restorer_param_shapes(cfg), gen)
    x = torch.rand(batch, cfg.range_dim, cfg.code_len, generator=gen).requires_grad_(True)
    mod = ref.Restorer(code_shape=(cfg.range_dim, cfg.code_len), soft=True, filters=cfg.dim, conv_type=1, expand=False, net_type="Linear")
    mod.load_state_dict(p)
    np.random.seed(seed)
    out = mod(x)
    assert out.shape == (batch, batch)
    np.random.seed(seed)
    noise = np.random.normal(0, 1, (batch, 1)).astype(np.float32)
    d_out = torch.randn(out.shape, generator=gen) * 0.1
    (out * d_out).sum().backward()
    store[prefix + "meta"] = np.array([seed, batch], dtype=np.int64)
    store[prefix + "x"] = x.detach().numpy(); store[prefix + "noise"] = noise; store[prefix + "out"] = out.detach().numpy()
    store[prefix + "d_out"] = d_out.numpy(); store[prefix + "d_x"] = x.grad.numpy()
    for k, v in mod.named_parameters():
        store[prefix + "grad." + k] = np.zeros(0, dtype=np.float32) if v.grad is None else v.grad.numpy()


def main():
    cfg = orc.PathConfig()
    store = {}
    for seed, batch in ((0, 5), (1, 48)):
        run_soft_case(store, f"soft.s{seed}.b{batch}.", seed, batch, cfg)
    for kind in ("res", "cls"):
        for seed, batch, training in ((0, 4, True), (1, 64, True), (2, 200, True), (3, 64, False)):
            run_case(store, f"{kind}.s{seed}.b{batch}.t{int(training)}.", kind, seed, batch, training, cfg)
    path = os.path.join(HERE, "iins_golden_convheads.npz")
    np.savez_compressed(path, **store)
    print("wrote", path, os.path.getsize(path), "bytes,", len(store), "arrays")


if __name__ == "__main__":
    main()
