"""Supervised training loop with the reference's signature (train.py:26):

    train_gem(opt, device, tensor, result_path, model_path, dataloader, val_dataloader, optimizer, network, data_raw)

``network`` is an ``iins_vae_b200.models.EMNet`` (Encoder -> Classifier + Restorer; the class is missing from the
reference, run.py:59-62).  The loop body -- CE(label_est, label) + L1(err_est, err), Adam step, running
RMSE / MAE / accuracy (train.py:82-115) -- runs in the fused engine; the Adam hyper-parameters are taken from
the torch ``optimizer`` the reference's caller builds (run.py:92-96), whose parameters are the network's, so
after the call the optimizer and the caller see the trained weights.  Validation every 20 epochs through
``test_gem`` (train.py:143-156), checkpoints as ``Network_%d.pth`` (:135-140).
"""
import datetime
import logging
import os
import sys
import time

import torch

from .engine import SemiTrainEngine
from .models import weights_init_normal


def train_gem(opt, device, tensor, result_path, model_path, dataloader, val_dataloader, optimizer, network, data_raw=None):
    logging.basicConfig(filename=os.path.join(result_path, "training_log.log"), level=logging.INFO)
    logging.info("Started")
    if opt.epoch != 0:
        network.load_state_dict(torch.load(os.path.join(model_path, "Network_%d.pth" % opt.epoch)))
    else:
        network.apply(weights_init_normal)
    g = optimizer.param_groups[0]
    engines = {}
    prev_time = time.time()
    history = []
    for epoch in range(opt.epoch, opt.n_epochs):
        rmse_sum = abs_sum = acc_sum = 0.0
        n_log = 0
        start_time = time.time()
        for i, batch in enumerate(dataloader):
            cir, err, label = batch["CIR"], batch["Err"], batch["Label"]
            B = cir.shape[0]
            eng = engines.get(B)
            if eng is None:
                eng = engines[B] = SemiTrainEngine(network.encoder, None, network.restorer, network.classifier, batch_size=B,
                                                   cir_len=cir.shape[1], lr=g["lr"], betas=tuple(g["betas"]), eps=g["eps"],
                                                   mode="supervised", shared_state=next(iter(engines.values()), None))
            eng.step(cir, err, label)
            if i % getattr(opt, "log_every", 1) == 0:
                t = eng.loss_terms()
                n_log += 1
                rmse_sum += t["rmse"]; abs_sum += t["mae"]; acc_sum += t["accuracy"]
                batches_done = epoch * len(dataloader) + i
                left = datetime.timedelta(seconds=(opt.n_epochs * len(dataloader) - batches_done) * (time.time() - prev_time) / max(batches_done, 1))
                line = ("\r[Data Env: %s] [Model Type: Identifier%s_Regressor%s] [Epoch: %d/%d] [Batch: %d/%d] "
                        "[Total Loss: %f, Idy Loss: %f, Reg Loss: %f] [Error: rmse %f, abs %f, accuracy %f] [Train Time: %f, ETA: %s]"
                        % (opt.dataset_env, opt.identifier_type, opt.regressor_type, epoch, opt.n_epochs, i, len(dataloader),
                           t["loss"], t["loss_env"], t["loss_res"], rmse_sum / n_log, abs_sum / n_log, acc_sum / n_log,
                           (time.time() - start_time) / (i + 1), left))
                sys.stdout.write(line)
                logging.info(line)
                history.append(t)
        if opt.checkpoint_interval != -1 and epoch % opt.checkpoint_interval == 0:
            torch.save(network.state_dict(), os.path.join(model_path, "Network_%d.pth" % epoch))
        if val_dataloader is not None and epoch % 20 == 0 and epoch != 0:
            from .test import test_gem
            test_gem(opt, device, tensor, result_path, model_path, val_dataloader, network, epoch, data_raw)
    return history
