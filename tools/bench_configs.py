"""Numbers for the other BASELINE.json configs (the driver's bench.py line is configs[1]):
  configs[2]  semi-supervised train step, bf16 tensor-core mode, batch 8192
  configs[4]  test.py inference (Encoder -> Restorer + Classifier + metrics) over N synthetic CIR windows
Prints one JSON line per config.  Single GPU; the multi-GPU variants shard the batch / the windows with no
data-path collective (inference) or one gradient all-reduce (training, see bench.py)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import iins_vae_b200
from oracle import iins_oracle as orc
from iins_vae_b200 import models as M
from iins_vae_b200.engine import SemiTrainEngine, InferenceEngine


def mods(cfg, seed=1234):
    pe, pd, pr, pc = orc.init_all(cfg, seed)
    Enc = M.Encoder(1, cfg.dim, cfg.n_residual, cfg.n_downsample, cfg.env_dim, cfg.range_dim)
    Dec = M.Decoder(1, cfg.dim, cfg.n_residual, cfg.n_downsample, cfg.env_dim, cfg.cir_len, cfg.range_dim)
    Res = M.Restorer((cfg.range_dim, cfg.code_len)); Cls = M.Classifier(cfg.env_dim, cfg.num_classes)
    for m, p in ((Enc, pe), (Dec, pd), (Res, pr), (Cls, pc)):
        m.load_state_dict(p); m.cuda()
    return Enc, Dec, Res, Cls


def timed(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    cfg = orc.PathConfig()
    out = []
    # ---- configs[2]: bf16, B=8192, semi step with the supervision mask
    for mode, B in (("bf16", 8192), ("fp32", 8192)):
        iins_vae_b200.set_compute_mode(mode)
        Enc, Dec, Res, Cls = mods(cfg)
        eng = SemiTrainEngine(Enc, Dec, Res, Cls, batch_size=B)
        batches = [tuple(t.cuda() for t in orc.synthetic_batch(cfg, B, 10 + j)) for j in range(4)]
        rng = np.random.RandomState(1234)
        masks = [orc.supervision_mask(rng, 0.1) for _ in range(64)]
        it = [0]
        def step():
            i = it[0]; it[0] += 1
            eng.step(*batches[i % 4], supervised=bool(masks[i % 64]))
        for sup in (True, False):
            eng.step(*batches[0], supervised=sup)
        ms = timed(step, 20, 5)
        out.append({"config": f"semi-supervised train step, {mode}, batch {B}, 1xB200", "ms_per_step": ms,
                    "samples_per_s": B / ms * 1e3, "final_loss": eng.loss_terms()["loss"]})
    # ---- configs[4]: inference over N windows
    iins_vae_b200.set_compute_mode("fp32")
    Enc, Dec, Res, Cls = mods(cfg)
    Bi = 65536
    n_windows = int(os.environ.get("IINS_INFER_WINDOWS", 1_250_000))       # one GPU's share of 10M windows on 8 GPUs
    n_batches = n_windows // Bi
    eng = InferenceEngine(Enc, Res, Cls, batch_size=Bi)
    g = torch.Generator(device="cuda").manual_seed(1)
    data = [torch.randn(Bi, cfg.cir_len, device="cuda", generator=g) for _ in range(4)]
    err = torch.rand(Bi, 1, device="cuda") * 0.3
    lab = torch.randint(0, cfg.num_classes, (Bi, 1), device="cuda").float()
    it = [0]
    def infer():
        i = it[0]; it[0] += 1
        eng.run(data[i % 4], err, lab)
    ms = timed(infer, n_batches, 3)
    out.append({"config": f"test.py inference (Enc+Res+Cls+metrics), fp32, {n_batches * Bi} windows in batches of {Bi}, 1xB200",
                "ms_per_batch": ms, "windows_per_s": Bi / ms * 1e3, "total_s": ms * n_batches / 1e3})
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()
