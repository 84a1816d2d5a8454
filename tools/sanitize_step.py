"""One supervised + one unsupervised train step and one inference pass at small batches (ragged and multi-tile) for
compute-sanitizer:  compute-sanitizer --tool memcheck python tools/sanitize_step.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from iins_vae_b200 import models as M
from iins_vae_b200.data import SyntheticCIR
from iins_vae_b200.engine import SemiTrainEngine, InferenceEngine

torch.manual_seed(0)
for B in (3, 160):
    Enc, Dec = M.Encoder(1, 4, 3, 4, 16, 2).cuda(), M.Decoder(1, 4, 3, 4, 16, 157, 2).cuda()
    Res, Cls = M.Restorer((2, 8)).cuda(), M.Classifier(16, 5).cuda()
    eng = SemiTrainEngine(Enc, Dec, Res, Cls, batch_size=B, use_graph=False)
    for batch, sup in zip(SyntheticCIR(2 * B, B, 157, 5, seed=1, pin=False), (True, False)):
        eng.step(batch["CIR"], batch["Err"], batch["Label"], supervised=sup)
    inf = InferenceEngine(Enc, Res, Cls, batch_size=B, use_graph=False)
    for batch in SyntheticCIR(B, B, 157, 5, seed=2, pin=False):
        inf.run(batch["CIR"], batch["Err"], batch["Label"])
    torch.cuda.synchronize()
    print("B", B, "loss", eng.loss_terms()["loss"])
print("done")
