"""On-device check of the fused residual-trunk kernels against the layer-by-layer path (same library, IINS_FUSED_TRUNK=0):
forward tensors of one encoder + decoder pass, both compute modes, whole and ragged batches."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(fused: str, out: str):
    env = dict(os.environ, IINS_FUSED_TRUNK=fused, IINS_FUSED_TRUNK_BWD=fused, IINS_TRUNK_TMAP=os.environ.get("IINS_TRUNK_TMAP", "1"))
    subprocess.run([sys.executable, __file__, "--child", out], env=env, check=True)


def child(out):
    import torch
    import iins_vae_b200
    from oracle import iins_oracle as orc
    from tests.test_gpu_parity import _mods
    from iins_vae_b200.engine import SemiTrainEngine
    res = {}
    cfg = orc.PathConfig()
    for mode in ("fp32", "bf16"):
        iins_vae_b200.set_compute_mode(mode)
        for B in (16, 37, 4096):
            mods, _ = _mods(cfg, 5)
            cir, err, label = orc.synthetic_batch(cfg, B, 7)
            eng = SemiTrainEngine(*mods, batch_size=B, use_graph=False)
            eng.step(cir, err, label, supervised=True, update=False)
            torch.cuda.synchronize()
            res[f"{mode}.{B}.rc"] = eng.rc.cpu()
            res[f"{mode}.{B}.xrec"] = eng.xrec.cpu()
            res[f"{mode}.{B}.loss"] = torch.tensor(eng.loss_terms()["loss"])
            for k, v in eng.named_grads().items():
                if k.startswith("enc.range_encoder") or k.startswith("dec.decoder"):
                    res[f"{mode}.{B}.g.{k}"] = v.cpu().clone()
    torch.save(res, out)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        child(sys.argv[2])
        sys.exit(0)
    import torch
    run("0", "/tmp/trunk_ref.pt")
    run("1", "/tmp/trunk_fused.pt")
    a, b = torch.load("/tmp/trunk_ref.pt"), torch.load("/tmp/trunk_fused.pt")
    bad = 0
    from oracle import iins_oracle as orc
    for k in a:
        if ".g." in k and orc.grad_is_structurally_zero(k.split(".g.")[1]):
            continue                        # conv biases in front of an InstanceNorm: the true gradient is exactly 0 (rounding noise only)
        d = float((a[k].double() - b[k].double()).norm())
        sc = float(a[k].double().norm()) + 1e-30
        # the two paths differ only in summation order / kink decisions: fp32-grade 2e-3 rel-L2 (a flipped ReLU kink at B = 16..37
        # moves a tensor by O(1/B)), bf16 operands 5e-2
        flag = "" if d <= (2e-3 if k.startswith("fp32.4096") else 5e-2) * sc else "  <<< MISMATCH"
        bad += bool(flag)
        if flag or ".g." not in k or "block.1.weight" in k or "mlp.model.4" in k or "model.2.weight" in k:
            print(f"{k:60s} rel-L2 diff {d / sc:.3e} (norm {sc:.3e}){flag}")
    print("TRUNK_CHECK", "FAILED" if bad else "PASSED")
    sys.exit(1 if bad else 0)
