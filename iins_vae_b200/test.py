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
    # Device-side accumulation: the per-batch metrics are summed in a device vector and the output
    # arrays are written into preallocated device buffers; the host synchronises ONCE, after the last batch (the
    # reference reads several scalars per batch and grows numpy arrays with vstack, test.py:76-107).
    acc = None                      # [sum of per-batch rmse, sum of per-batch mae, sum of per-batch accuracy]
    total = None                    # number of windows, when the loader can tell without being iterated
    ds = getattr(dataloader, "dataset", None)
    if ds is not None and hasattr(ds, "__len__"):
        total = len(ds)
    elif hasattr(dataloader, "batch_size") and hasattr(dataloader, "__len__"):
        total = len(dataloader) * int(dataloader.batch_size)
    err_buf = pred_buf = lat_buf = None
    err_all, pred_all, latent_all = [], [], []
    start_time = time.time()
    n = 0
    off = 0
    for i, batch in enumerate(dataloader):
        cir, err, label = batch["CIR"], batch["Err"], batch["Label"]
        B = cir.shape[0]
        eng = engines.get(B)
        if eng is None:
            eng = engines[B] = InferenceEngine(network.encoder, network.restorer, network.classifier, batch_size=B,
                                               cir_len=cir.shape[1])
        err_est, pred, out = eng.run(cir, err, label)
        if acc is None:
            acc = torch.zeros(3, dtype=torch.float64, device=out.device)
            if total is not None:
                err_buf = torch.empty(total, 1, device=out.device); pred_buf = torch.empty(total, dtype=torch.int32, device=out.device)
                lat_buf = torch.empty(total, eng.cat.shape[1], device=out.device)
        acc[0] += out[4].clamp_min(0.0).sqrt()               # mean of per-batch RMSEs, like the reference (test.py:76-85)
        acc[1] += out[1]
        acc[2] += out[5] / B
        n += 1
        if err_buf is not None and off + B <= total:
            err_buf[off:off + B].copy_(err_est); pred_buf[off:off + B].copy_(pred); lat_buf[off:off + B].copy_(eng.cat)
        else:
            err_all.append(err_est.clone()); pred_all.append(pred.clone()); latent_all.append(eng.cat.clone())
        off += B
    sums = acc.tolist() if acc is not None else [0.0, 0.0, 0.0]          # the only host synchronisation
    rmse_sum, abs_sum, acc_sum = sums
    if err_buf is not None and not err_all:
        err_cat, pred_cat, lat_cat = err_buf[:off], pred_buf[:off], lat_buf[:off]
    else:
        err_cat = torch.cat(err_all) if err_all else None
        pred_cat = torch.cat(pred_all) if pred_all else None
        lat_cat = torch.cat(latent_all) if latent_all else None
    time_avg = (time.time() - start_time) / max(n, 1) / 500           # test.py:78 divides by the hard-coded 500
    res = dict(rmse=rmse_sum / max(n, 1), abs=abs_sum / max(n, 1), accuracy=acc_sum / max(n, 1), time=time_avg,
               err_est=err_cat, pred=pred_cat, env_latent=lat_cat)
    line = "[Data Env: %s] [Epoch: %d] [Error: rmse %f, abs %f, accuracy %f] [Test Time: %f]" % (
        opt.dataset_env, epoch, res["rmse"], res["abs"], res["accuracy"], time_avg)
    print(line)
    logging.info(line)
    return res
