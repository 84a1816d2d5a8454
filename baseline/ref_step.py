"""Drive the UNMODIFIED reference modules (baseline/_ref/models.py, installed by baseline/install_ref.py) through the
reference's own training step.  train_semi.py itself cannot run (it reads five undeclared options and imports files
with syntax errors: SURVEY.md 8(c)), so its loop body :183-228 is restated here around the reference's nn.Modules,
criteria and torch.optim.Adam -- none of this repository's kernels, modules or engine is on this path.

Used by `bench.py --impl reference` (CPU, all host threads), by bench.py's `cpu_baseline`, and by its
`gpu_eager_reference` leg (the same modules in eager PyTorch-CUDA: "the kernel to beat on the same box").
"""
import importlib.util
import itertools
import json
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_MODELS = os.path.join(HERE, "_ref", "models.py")


def available() -> bool:
    return os.path.exists(REF_MODELS)


def manifest():
    p = os.path.join(HERE, "_ref", "MANIFEST.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return {}


def load_reference_models():
    spec = importlib.util.spec_from_file_location("iins_reference_models", REF_MODELS)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class ReferenceTrainer:
    """train_semi.py:77-82 (modules), :104-107 (init), :111-122 (loss weights, Adam), :183-228 (step)."""

    def __init__(self, device, dim=4, n_residual=3, n_downsample=4, env_dim=16, range_dim=2, num_classes=5, cir_len=157,
                 lr=1e-4, b1=0.5, b2=0.999, seed=1234):
        ref = load_reference_models()
        torch.manual_seed(seed)
        self.device = torch.device(device)
        self.Enc = ref.Encoder(conv_type=1, dim=dim, n_downsample=n_downsample, n_residual=n_residual, style_dim=env_dim,
                               out_dim=range_dim, expand=False).to(self.device)
        self.Dec = ref.Decoder(conv_type=1, dim=dim, n_upsample=n_downsample, n_residual=n_residual, style_dim=env_dim,
                               in_dim=cir_len, out_dim=range_dim, expand=False).to(self.device)
        self.Res = ref.Restorer(code_shape=(range_dim, 128 // 2 ** n_downsample), soft=False, filters=dim, conv_type=1,
                                expand=False, net_type="Linear").to(self.device)
        self.Cls = ref.Classifier(env_dim=env_dim, num_classes=num_classes, filters=16, net_type="Linear").to(self.device)
        for m in (self.Enc, self.Dec, self.Res, self.Cls):
            m.apply(ref.weights_init_normal)
        self.criterion_recon = torch.nn.L1Loss().to(self.device)
        self.criterion_code = torch.nn.CrossEntropyLoss().to(self.device)
        self.optimizer = torch.optim.Adam(itertools.chain(self.Enc.parameters(), self.Dec.parameters(), self.Res.parameters(),
                                                          self.Cls.parameters()), lr=lr, betas=(b1, b2))
        self.lambda_ae, self.lambda_res, self.lambda_range, self.lambda_env = 1, 10, 1, 1

    def step(self, cir_gt, err_gt, label_gt, mask: int):
        """One iteration of the loop body; returns the loss tensor (no .item(): the caller decides when to sync)."""
        cir_gt = cir_gt.to(self.device)
        err_gt = err_gt.to(self.device)
        label_gt = label_gt.to(device=self.device, dtype=torch.int64)
        self.optimizer.zero_grad()
        range_code, env_code, env_code_rv, kl_div = self.Enc(cir_gt)
        cir_gen = self.Dec(range_code, env_code)
        err_fake = self.Res(range_code)
        label_fake = self.Cls(env_code)
        loss_ae = self.lambda_ae * self.criterion_recon(cir_gt, cir_gen)
        loss_range = self.lambda_range * kl_div
        if mask == 0:
            loss = loss_ae + loss_range
            loss.backward()
            self.optimizer.step()
            return loss
        label_gt = label_gt.squeeze()
        loss_res = self.lambda_res * self.criterion_recon(err_gt, err_fake)
        loss_env = self.lambda_env * self.criterion_code(label_fake, label_gt)       # dataset_env == 'room_full'
        loss = loss_ae + loss_range + loss_res + loss_env
        loss.backward()
        self.optimizer.step()
        return loss


def synthetic_batches(n, batch, cir_len=157, num_classes=5, seed=1234):
    """Host batches of the zenodo loader's shape (dataset.py:118-133), same statistics as iins_vae_b200.data.SyntheticCIR."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n):
        cir = torch.randn(batch, cir_len, generator=g)
        err = (torch.randn(batch, 1, generator=g) * 0.15).abs().clamp_(0, 1)
        label = torch.randint(0, num_classes, (batch, 1), generator=g).float()
        out.append((cir, err, label))
    return out


def mask_stream(rate=0.1, seed=1234):
    rng = np.random.RandomState(seed)
    while True:
        yield 0 if rng.randn(1)[0] > rate else 1                                  # train_semi.py:203
