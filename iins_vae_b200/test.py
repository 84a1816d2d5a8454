"""Evaluation loop with the reference's signature (test.py:26):

    test_gem(opt, device, tensor, result_path, model_path, dataloader, network, epoch, data_raw)

Covers the hot part of the reference's function (test.py:55-85): no-grad ``network(cir)`` and the per-batch
RMSE / MAE / argmax-accuracy, averaged over batches like the reference (mean of per-batch values).  The tail
of the reference function (UMAP, CDF plots, SVM baselines, savemat: test.py:89-146) is outside the hot path
 -- the arrays it would consume are returned instead.
"""
import logging
import os
import time

import torch

from .engine import InferenceEngine


def test_gem(opt, device, tensor, result_path, model_path, dataloader, network, epoch, data_raw=None):
    logging.basicConfig(filename=os.path.join(result_path, "val_log.log"), level=logging.INFO)
    logging.info("Started")
    if epoch != 0:
        network.load_state_dict(torch.load(os.path.join(model_path, "Network_%d.pth" % epoch)))
        network.eval()
    else:
        print("No saved models in dirs.")
    engines = {}
    rmse_sum = abs_sum = acc_sum = 0.0
    err_all, pred_all, latent_all = [], [], []
    start_time = time.time()
    n = 0
    for i, batch in enumerate(dataloader):
        cir, err, label = batch["CIR"], batch["Err"], batch["Label"]
        B = cir.shape[0]
        eng = engines.get(B)
        if eng is None:
            eng = engines[B] = InferenceEngine(network.encoder, network.restorer, network.classifier, batch_size=B,
                                               cir_len=cir.shape[1])
        err_est, pred, out = eng.run(cir, err, label)
        o = out.tolist()                                      # one sync per batch (the reference has several)
        rmse_sum += max(o[4], 0.0) ** 0.5
        abs_sum += o[1]
        acc_sum += o[5] / B
        n += 1
        err_all.append(err_est.clone())
        pred_all.append(pred.clone())
        latent_all.append(eng.cat.clone())
    time_avg = (time.time() - start_time) / max(n, 1) / 500           # test.py:78 divides by the hard-coded 500
    res = dict(rmse=rmse_sum / max(n, 1), abs=abs_sum / max(n, 1), accuracy=acc_sum / max(n, 1), time=time_avg,
               err_est=torch.cat(err_all) if err_all else None, pred=torch.cat(pred_all) if pred_all else None,
               env_latent=torch.cat(latent_all) if latent_all else None)
    line = "[Data Env: %s] [Epoch: %d] [Error: rmse %f, abs %f, accuracy %f] [Test Time: %f]" % (
        opt.dataset_env, epoch, res["rmse"], res["abs"], res["accuracy"], time_avg)
    print(line)
    logging.info(line)
    return res
