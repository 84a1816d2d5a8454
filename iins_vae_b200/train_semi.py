"""Semi-supervised training entry point -- the flow of the reference's ``train_semi.py`` (:24-286) on the
fused B200 engine.  ``python -m iins_vae_b200.train_semi --dataset_env room_full --synthetic 65536 --batch_size 4096``

Same options (utils.get_args + --supervision_rate), same module construction (train_semi.py:77-82), init
(:104-107), Adam hyper-parameters (:118-122), LambdaLR schedule (:126-128), per-batch supervision mask (:203),
loss weights (:111-114), log line (:257-268) and checkpoint names (:281-286).  Differences, all deliberate:
the loss / metric scalars stay on the device and are read every --log_every steps instead of 5x .item() per
step; the data comes from SyntheticCIR unless ``dataloader`` is passed to ``run``.
"""
import datetime
import logging
import os
import sys
import time

import torch

from . import set_compute_mode
from .data import SyntheticCIR
from .engine import SemiTrainEngine
from .models import Classifier, Decoder, Encoder, LambdaLR, Restorer, weights_init_normal
from .parallel import SupervisionMask, broadcast_parameters, init_distributed, shutdown_distributed
from .utils import get_args, num_classes_for


def label_offset_for(dataset_env: str) -> int:
    """train_semi.py:217-222 / :250-253: CrossEntropyLoss gets ``label_gt - 1`` and accuracy compares ``argmax + 1`` for every
    dataset_env except 'room_full' (whose labels are 0..4); the other environments carry labels 1..NC."""
    return 0 if dataset_env == "room_full" else 1


class _Lookahead:
    """Iterate a loader one batch ahead so that the next batch can be handed to ``SemiTrainEngine.prefetch``."""

    def __init__(self, it):
        self.it = iter(it)
        self.nxt = next(self.it, None)

    def __iter__(self):
        return self

    def __next__(self):
        if self.nxt is None:
            raise StopIteration
        cur, self.nxt = self.nxt, next(self.it, None)
        return cur

    def peek(self):
        return self.nxt


def build_modules(opt, device):
    """train_semi.py:43-82."""
    len_cir = 157
    opt.num_classes = num_classes_for(opt.dataset_env)
    opt.if_expand = False if opt.conv_type == 1 else True
    range_code_shape = (opt.range_dim, 128 // (2 ** opt.n_downsample))
    Enc = Encoder(conv_type=opt.conv_type, dim=opt.dim, n_downsample=opt.n_downsample, n_residual=opt.n_residual,
                  style_dim=opt.env_dim, out_dim=opt.range_dim, expand=opt.if_expand).to(device)
    Dec = Decoder(conv_type=opt.conv_type, dim=opt.dim, n_upsample=opt.n_downsample, n_residual=opt.n_residual,
                  style_dim=opt.env_dim, in_dim=len_cir, out_dim=opt.range_dim, expand=opt.if_expand).to(device)
    Res = Restorer(code_shape=range_code_shape, soft=False, filters=opt.dim, conv_type=opt.conv_type,
                   expand=opt.if_expand, net_type=opt.restorer_type).to(device)
    Cls = Classifier(env_dim=opt.env_dim, num_classes=opt.num_classes, filters=16, net_type=opt.classifier_type).to(device)
    return Enc, Dec, Res, Cls, len_cir


last_modules = None          # (Enc, Dec, Res, Cls) of the most recent run() in this process (for callers that continue with them)


def run(opt, dataloader=None, max_steps=None, quiet=False):
    """The training loop; under torchrun the process group is torn down in an orderly way at the end."""
    engines = {}
    try:
        return _run(opt, engines, dataloader, max_steps, quiet)
    finally:
        if int(os.environ.get("WORLD_SIZE", "1")) > 1:
            shutdown_distributed(engines.values())


def _run(opt, engines, dataloader=None, max_steps=None, quiet=False):
    rank, local, world, pg = init_distributed()
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    set_compute_mode(opt.compute_mode)
    Enc, Dec, Res, Cls, len_cir = build_modules(opt, device)
    global last_modules
    last_modules = (Enc, Dec, Res, Cls)
    tag = "%s_mode_%s/SEMI%f_AE%d_Res%s_Cls%s_Rdim%dEdim%d" % (opt.dataset_env, opt.mode, opt.supervision_rate, opt.conv_type,
                                                           opt.restorer_type, opt.classifier_type, opt.range_dim, opt.env_dim)
    model_path, result_path = os.path.join("saved_models_semi", tag), os.path.join("saved_results_semi", tag)
    if rank == 0:
        os.makedirs(model_path, exist_ok=True)
        os.makedirs(result_path, exist_ok=True)
        logging.basicConfig(filename=os.path.join(result_path, "train_log.log"), level=logging.INFO)
        logging.info("Started")
    if opt.epoch != 0:                                                     # train_semi.py:97-102
        for m, n in ((Enc, "Enc"), (Dec, "Dec"), (Res, "Res"), (Cls, "Cls")):
            m.load_state_dict(torch.load(os.path.join(model_path, "%s_%d.pth" % (n, opt.epoch))))
    else:
        for m in (Enc, Dec, Res, Cls):
            m.apply(weights_init_normal)
    broadcast_parameters((Enc, Dec, Res, Cls), pg)
    root = getattr(opt, "data_root", "./data/data_zenodo/dataset.pkl")                 # train_semi.py:133
    if dataloader is None and opt.synthetic <= 0 and os.path.exists(root):
        # the reference's own data flow (train_semi.py:137-147) on the pinned-ring pipeline: pickle -> split -> StandardScaler
        # fitted on the training CIRs -> shuffled batches gathered into pinned host buffers by a background thread
        from .dataset import PinnedBatchRing, UWBDataset, err_mitigation_dataset
        data_train, _, _, _ = err_mitigation_dataset(root=root, dataset_name=opt.dataset_name, dataset_env=opt.dataset_env,
                                                     split_factor=0.8, scaling=True, mode=opt.mode)
        if world > 1:                                                                  # every rank trains on its own slice
            data_train = tuple(a[rank::world] for a in data_train)
        dataloader = PinnedBatchRing(UWBDataset(data_train), opt.batch_size, shuffle=True, seed=1234 + rank)
    if dataloader is None:
        n = opt.synthetic if opt.synthetic > 0 else 16 * opt.batch_size
        dataloader = SyntheticCIR(n, opt.batch_size, len_cir, opt.num_classes, seed=1234 + rank,
                                  label_base=label_offset_for(opt.dataset_env))
    sched = LambdaLR(opt.n_epochs, opt.epoch, opt.decay_epoch)
    mask_stream = SupervisionMask(opt.supervision_rate, seed=1234)
    prev_time, steps_done = time.time(), 0
    last = {}
    for epoch in range(opt.epoch, opt.n_epochs):
        lr = opt.lr * sched.step(epoch - opt.epoch)                        # LambdaLR(optimizer, lr_lambda=...) semantics
        rmse_sum = abs_sum = acc_sum = 0.0
        n_sup = 0
        lookahead = _Lookahead(dataloader)
        pending = None
        for i, batch in enumerate(lookahead):
            cir, err, label = batch["CIR"], batch["Err"], batch["Label"]
            B = cir.shape[0]
            eng = engines.get(B)
            if eng is None:
                eng = engines[B] = SemiTrainEngine(Enc, Dec, Res, Cls, batch_size=B, cir_len=cir.shape[1], lr=lr,
                                                   betas=(opt.b1, opt.b2), mode="semi", process_group=pg,
                                                   shared_state=next(iter(engines.values()), None),
                                                   label_offset=label_offset_for(opt.dataset_env))
                opt_file = os.path.join(model_path, "Opt_%d.pth" % opt.epoch)
                if opt.epoch != 0 and len(engines) == 1 and os.path.exists(opt_file):
                    eng.load_optimizer_state_dict(torch.load(opt_file))   # exact resume (the reference restarts Adam)
            eng.set_lr(lr)
            supervised = bool(mask_stream())
            if pending is eng:                                             # this batch was prefetched during the previous step
                eng.step(supervised=supervised, prefetched=True)
            else:
                eng.step(cir, err, label, supervised=supervised)
            pending = None
            nxt = lookahead.peek()
            if nxt is not None and nxt["CIR"].shape[0] == B and nxt["CIR"].is_pinned():
                eng.prefetch(nxt["CIR"], nxt["Err"], nxt["Label"])        # H2D of the next batch overlaps this step
                pending = eng
            steps_done += 1
            if supervised and (i % opt.log_every == 0 or max_steps is not None):
                t = eng.loss_terms()                                       # the only host sync
                n_sup += 1
                rmse_sum += t["rmse"]; abs_sum += t["mae"]; acc_sum += t["accuracy"]
                last = t
                if rank == 0 and not quiet:
                    left = datetime.timedelta(seconds=(opt.n_epochs * len(dataloader) - steps_done) * (time.time() - prev_time) / max(steps_done, 1))
                    line = ("\r[Model Name: C%d_%s_semi%f] [Epoch: %d/%d] [Batch %d/%d] [RMSE: %F] [ABS ERROR: %F] [Accuracy: %f] "
                            "[Train Time: %f] [Total loss: %f] [Supervised loss: ae %f, kl %f] [Unsup loss: res %f, cls %f] ETA: %s"
                            % (opt.conv_type, opt.restorer_type, opt.supervision_rate, epoch, opt.n_epochs, i, len(dataloader),
                               rmse_sum / n_sup, abs_sum / n_sup, acc_sum / n_sup, (time.time() - prev_time) / max(steps_done, 1) / B,
                               t["loss"], t["loss_ae"], t["loss_range"], t["loss_res"], t["loss_env"], left))
                    sys.stdout.write(line)
                    logging.info(line)
            if max_steps is not None and steps_done >= max_steps:
                return last
        if rank == 0 and opt.checkpoint_interval != -1 and epoch % opt.checkpoint_interval == 0:
            for m, n in ((Enc, "Enc"), (Dec, "Dec"), (Res, "Res"), (Cls, "Cls")):
                torch.save(m.state_dict(), os.path.join(model_path, "%s_%d.pth" % (n, epoch)))
            if engines:                                                    # beyond the reference: Adam moments + step counters
                torch.save(next(iter(engines.values())).optimizer_state_dict(), os.path.join(model_path, "Opt_%d.pth" % epoch))
    return last


def main(argv=None):
    parser = get_args(None)
    parser.add_argument("--supervision_rate", type=float, default=0.1, help="Rate of labeled data to pure cir data.")
    parser.set_defaults(dataset_env="room_full")       # train_semi.py:46-63 has no 'nlos' branch
    opt = parser.parse_args(argv)
    print(opt)
    return run(opt)


if __name__ == "__main__":
    main()
