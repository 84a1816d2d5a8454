"""Fused train / inference engines: the whole step of train_semi.py:183-228 (or train.py:75-94) as a fixed
sequence of C-ABI calls on one CUDA stream -- no autograd graph, no per-step allocation, one flat
parameter / gradient / Adam-state buffer, optionally replayed from a CUDA graph.

The nn.Modules stay the owners of the parameters (their ``.data`` become views into the flat buffer), so
``state_dict()`` / ``torch.save`` keep working exactly like the reference's checkpoints
(train_semi.py:281-286).
"""
import ctypes as C

import numpy as np
import torch

from ._capi import IinsConfig, IinsHeadState, get_lib, ptr, ptr_array

LAMBDA_AE, LAMBDA_RES, LAMBDA_RANGE, LAMBDA_ENV = 1.0, 10.0, 1.0, 1.0      # train_semi.py:111-114


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


_SHARED_STREAMS = {}


def _shared_stream(device, role):
    """Side streams are shared by all engines of a device: the library keeps a helper-stream pair per caller stream."""
    key = (torch.device(device).index, role)
    if key not in _SHARED_STREAMS:
        _SHARED_STREAMS[key] = torch.cuda.Stream(device=device)
    return _SHARED_STREAMS[key]


def _capture_stream(device):
    """The stream a step is captured on = the stream of the step's critical path (forward chain, loss, data-gradient chain, Adam).
    IINS_MAIN_PRIORITY=-1 creates it with a higher stream priority than the library's helper streams (weight gradients, the
    env-encoder branch, the heads); measured on the B200: no effect on the step time (1.6986 vs 1.6993 ms: every kernel fills
    the machine, there is rarely a choice between pending thread blocks), so the default stays 0."""
    import os
    prio = int(os.environ.get("IINS_MAIN_PRIORITY", "0"))
    key = (torch.device(device).index, "capture", prio)
    if key not in _SHARED_STREAMS:
        _SHARED_STREAMS[key] = torch.cuda.Stream(device=device, priority=prio)
    return _SHARED_STREAMS[key]


def _warmup_stream(device):
    """The side stream on which a step is run once eagerly before it is captured."""
    return _shared_stream(device, "warmup")


class _Flat:
    """Flatten the parameters of several modules into one buffer (module order, named_parameters order)."""

    def __init__(self, modules, device):
        self.params, self.spans = [], []
        off = 0
        for m in modules:
            begin = off
            for p in m.parameters():
                self.params.append(p)
                off += p.numel()
            self.spans.append((begin, off))
        self.total = off
        self.flat = torch.empty(off, dtype=torch.float32, device=device)
        self.grad = torch.zeros(off, dtype=torch.float32, device=device)
        self.exp_avg = torch.zeros(off, dtype=torch.float32, device=device)
        self.exp_avg_sq = torch.zeros(off, dtype=torch.float32, device=device)
        o = 0
        self.grad_views = []
        for p in self.params:
            n = p.numel()
            self.flat[o:o + n].copy_(p.data.reshape(-1).to(device=device, dtype=torch.float32))
            p.data = self.flat[o:o + n].view(p.shape)
            self.grad_views.append(self.grad[o:o + n].view(p.shape))
            o += n


class SemiTrainEngine:
    """One object = the reference's (Enc, Dec, Res, Cls, Adam) training state, stepped by fused kernels.

    mode="semi"       : loss_ae + kl (+ 10*L1(err) + CE when the batch is supervised)  -- train_semi.py:199-225
    mode="supervised" : CE + L1(err) on Encoder -> (Classifier, Restorer), no decoder -- train.py:82-91
    """

    def __init__(self, Enc, Dec, Res, Cls, batch_size, cir_len=157, lr=1e-4, betas=(0.5, 0.999), eps=1e-8,
                 mode="semi", use_graph=True, process_group=None, device=None, shared_state=None, label_offset=0,
                 overlap_allreduce=True):
        """``shared_state``: another engine over the SAME modules (e.g. for a different batch size, the last
        ragged batch of an epoch): the flat parameter / gradient / Adam buffers and step counters are shared.
        ``label_offset``: class index = label - label_offset (train_semi.py:217-222: 1 for every dataset_env except
        'room_full', whose labels are already 0-based).  ``overlap_allreduce``: reduce the Dec/Res/Cls gradient bucket on a
        communication stream while the encoder backward runs; False = one all-reduce after the backward."""
        self.lib = get_lib()
        self.mode = mode
        self.device = torch.device(device if device is not None else torch.cuda.current_device())
        if self.device.type != "cuda":
            raise RuntimeError("iins_vae_b200 engines run on CUDA only")
        self.Enc, self.Dec, self.Res, self.Cls = Enc, Dec, Res, Cls
        mods = [Enc] + ([Dec] if mode == "semi" else []) + [Res, Cls]
        self.flat = shared_state.flat if shared_state is not None else _Flat(mods, self.device)
        o = Enc.opts
        self.B, self.L = int(batch_size), int(cir_len)
        self.E, self.R, self.NC = o["env_dim"], o["range_dim"], Cls.num_classes
        self.cfg = IinsConfig(self.B, self.L, o["dim"], o["n_residual"], o["n_downsample"], self.E, self.R, self.NC,
                              Cls.filters)
        self.lib.check(self.lib.iins_validate_config(self.cfg), "engine config")
        self.betas, self.eps = betas, eps
        self.label_offset = int(label_offset)
        self.overlap_allreduce = bool(overlap_allreduce)
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
        self.rank = torch.distributed.get_rank(process_group) if process_group is not None else 0
        dev, B = self.device, self.B
        f = lambda *shape: torch.zeros(*shape, dtype=torch.float32, device=dev)
        # static I/O
        self.cir, self.err, self.label = f(B, self.L), f(B, 1), f(B, 1)
        self.rc, self.cat, self.kl = f(B, self.R, 128 >> o["n_downsample"]), f(B, self.E), f(1)
        self.xrec, self.err_est, self.logits = f(B, self.L), f(B, 1), f(B, self.NC)
        self.out = f(8)
        self.pred = torch.zeros(B, dtype=torch.int32, device=dev)
        self.d_xrec, self.d_err, self.d_logits = f(B, self.L), f(B, 1), f(B, self.NC)
        self.d_rc, self.d_cat = torch.zeros_like(self.rc), torch.zeros_like(self.cat)
        # the two heads run on their own stream next to the decoder (independent consumers of the encoder outputs):
        # their range_code / env_code gradients land in separate buffers and are summed once both sides are done
        self.d_rc_heads, self.d_cat_heads = torch.zeros_like(self.rc), torch.zeros_like(self.cat)
        self.head_stream = _shared_stream(self.device, "heads")     # shared by all engines of the device (one host thread drives them)
        self.comm_stream = _shared_stream(self.device, "comm")      # gradient all-reduce next to the encoder backward
        self.d_kl = torch.full((1,), LAMBDA_RANGE, dtype=torch.float32, device=dev)
        if shared_state is not None:
            self.lr, self.steps = shared_state.lr, shared_state.steps
        else:
            self.lr = torch.full((1,), float(lr), dtype=torch.float32, device=dev)
            self.steps = torch.zeros(8, dtype=torch.int32, device=dev)
        lib = self.lib
        names = ("encoder", "decoder", "restorer", "classifier")
        # Conv1d heads (net_type='Conv1d') have their own entry points, workspaces and a BatchNorm state
        self.conv_head = {"restorer": getattr(Res, "net_type", "Linear") == "Conv1d",
                          "classifier": getattr(Cls, "net_type", "Linear") == "Conv1d"}
        q = lambda m, what: f"iins_{m}_conv_{what}_floats" if self.conv_head.get(m) else f"iins_{m}_{what}_floats"
        self.ws = {m: f(int(getattr(lib, q(m, "ws"))(self.cfg)) + 16) for m in names}
        self.scratch = {m: f(int(getattr(lib, q(m, "scratch"))(self.cfg)) + 16) for m in names}
        self.head_state = {}
        for m, mod in (("restorer", Res), ("classifier", Cls)):
            if self.conv_head[m]:
                bn = getattr(getattr(mod, m).conv_blocks, "6")
                stats = torch.zeros(4 * bn.num_features, dtype=torch.float64, device=dev)
                self.head_state[m] = dict(bn=bn, stats=stats, C=bn.num_features, seed=mod._seed)
        # pointer tables per module
        self._tables(mods)
        self._graphs = {}
        # NCCL all-reduce is capturable: the data-parallel step is replayed from a CUDA graph as well
        self.use_graph = use_graph
        self.n_steps = 0

    # ------------------------------------------------------------------------------------------------
    def _tables(self, mods):
        fl = self.flat
        self.ptab, self.gtab, self.mod_span = {}, {}, {}
        i = 0
        for m, name in zip(mods, ["enc"] + (["dec"] if self.mode == "semi" else []) + ["res", "cls"]):
            n = len(list(m.parameters()))
            self.ptab[name] = ptr_array(fl.params[i:i + n])
            self.gtab[name] = ptr_array(fl.grad_views[i:i + n])
            i += n
        spans = dict(zip(["enc"] + (["dec"] if self.mode == "semi" else []) + ["res", "cls"], fl.spans))
        self.spans = spans
        # Adam groups (half-open element ranges): always-on [Enc (+Dec)], Res without linear_layer2, Cls.
        res_b, res_e = spans["res"]
        # restorer.linear_layer2.* (the soft head, unused when soft=False) are the trailing parameters of the Restorer
        res_used = res_e - sum(p.numel() for n, p in self.Res.named_parameters() if "linear_layer2" in n)
        always_end = spans["dec"][1] if self.mode == "semi" else spans["enc"][1]
        self.groups = [(0, always_end), (res_b, res_used), spans["cls"]]
        self._gb = (C.c_int64 * 3)(*[g[0] for g in self.groups])
        self._ge = (C.c_int64 * 3)(*[g[1] for g in self.groups])

    def set_lr(self, lr: float):
        self.lr.fill_(float(lr))

    # ------------------------------------------------------------------------------------------------
    concurrent = True       # False: one stream, no overlap (per-kernel event timing); see set_concurrency()

    def set_concurrency(self, enable: bool):
        """Overlap of independent work on helper streams (heads next to the decoder here; weight gradients and the env
        encoder inside the library).  Disable to time individual kernels; captured graphs keep the setting they had."""
        self.concurrent = bool(enable)
        self.lib.check(self.lib.iins_set_stream_concurrency(int(enable)), "set_stream_concurrency")

    def _heads_concurrent(self, supervised: bool) -> bool:
        return self.concurrent and self.mode == "semi" and supervised

    def _forward(self, supervised: bool):
        lib, cfg, st = self.lib, self.cfg, _stream()
        lib.check(lib.iins_encoder_forward(cfg, self.ptab["enc"], ptr(self.cir), None, 0, 0, ptr(self.rc), ptr(self.cat),
                                           None, ptr(self.kl), ptr(self.ws["encoder"]), st), "encoder forward")

        def heads(hst):
            for m, tab, src, dst in (("restorer", "res", self.rc, self.err_est), ("classifier", "cls", self.cat, self.logits)):
                if self.conv_head[m]:
                    self._conv_head_call(m, "forward", hst, lambda st: getattr(lib, f"iins_{m}_conv_forward")(
                        cfg, self.ptab[tab], ptr(src), ptr(dst), ptr(self.ws[m]), C.byref(st), hst), 0)
                else:
                    lib.check(getattr(lib, f"iins_{m}_forward")(cfg, self.ptab[tab], ptr(src), ptr(dst), ptr(self.ws[m]), hst), f"{m} forward")

        main = torch.cuda.current_stream()
        if self._heads_concurrent(supervised):
            self.head_stream.wait_stream(main)
            with torch.cuda.stream(self.head_stream):
                heads(_stream())
        if self.mode == "semi":
            lib.check(lib.iins_decoder_forward(cfg, self.ptab["dec"], ptr(self.rc), ptr(self.cat), ptr(self.xrec),
                                               ptr(self.ws["decoder"]), st), "decoder forward")
        if self._heads_concurrent(supervised):
            main.wait_stream(self.head_stream)
        elif supervised:
            heads(st)

    def _loss(self, supervised: bool):
        lib, semi = self.lib, self.mode == "semi"
        lam_res = LAMBDA_RES if semi else 1.0
        lib.check(lib.iins_loss_forward_backward(
            self.B, self.L, self.NC, ptr(self.cir) if semi else None, ptr(self.xrec) if semi else None,
            ptr(self.err) if supervised else None, ptr(self.err_est) if supervised else None,
            ptr(self.logits) if supervised else None, ptr(self.label) if supervised else None, None, self.label_offset,
            LAMBDA_AE, lam_res, LAMBDA_ENV, ptr(self.out), ptr(self.d_xrec) if semi else None,
            ptr(self.d_err) if supervised else None, ptr(self.d_logits) if supervised else None, ptr(self.pred),
            _stream()), "loss")

    def _backward(self, supervised: bool):
        lib, cfg, st, semi = self.lib, self.cfg, _stream(), self.mode == "semi"
        # the gradient buffer is zero here: the fused Adam zeroes what it consumes (iins_adam_step zero_grads), and a
        # gradients-only step (update=False) is followed by an explicit clear in step()
        conc = self._heads_concurrent(supervised)
        main = torch.cuda.current_stream()

        def heads(hst, d_rc, d_cat, acc):
            for m, tab, src, dout, din in (("restorer", "res", self.rc, self.d_err, d_rc), ("classifier", "cls", self.cat, self.d_logits, d_cat)):
                if self.conv_head[m]:
                    self._conv_head_call(m, "backward", hst, lambda st: getattr(lib, f"iins_{m}_conv_backward")(
                        cfg, self.ptab[tab], ptr(src), ptr(self.ws[m]), ptr(dout), self.gtab[tab], ptr(din), acc, ptr(self.scratch[m]),
                        C.byref(st), hst), 2)
                else:
                    lib.check(getattr(lib, f"iins_{m}_backward")(cfg, self.ptab[tab], ptr(src), ptr(self.ws[m]), ptr(dout), self.gtab[tab],
                                                                 ptr(din), acc, ptr(self.scratch[m]), hst), f"{m} backward")

        if conc:
            self.head_stream.wait_stream(main)
            with torch.cuda.stream(self.head_stream):
                heads(_stream(), self.d_rc_heads, self.d_cat_heads, 0)
        acc = 0
        if semi:
            lib.check(lib.iins_decoder_backward(cfg, self.ptab["dec"], ptr(self.rc), ptr(self.cat), ptr(self.ws["decoder"]),
                                                ptr(self.d_xrec), self.gtab["dec"], ptr(self.d_rc), ptr(self.d_cat), 0,
                                                ptr(self.scratch["decoder"]), st), "decoder backward")
            acc = 1
        if conc:
            main.wait_stream(self.head_stream)
            lib.check(lib.iins_accumulate2(ptr(self.d_rc), ptr(self.d_rc_heads), self.d_rc.numel(), ptr(self.d_cat),
                                           ptr(self.d_cat_heads), self.d_cat.numel(), st), "accumulate head gradients")
        elif supervised:
            heads(st, self.d_rc, self.d_cat, acc)
        self._allreduce_late_buckets(supervised)
        lib.check(lib.iins_encoder_backward(cfg, self.ptab["enc"], None, 0, 0, ptr(self.rc), ptr(self.cat),
                                            ptr(self.ws["encoder"]), ptr(self.d_rc), ptr(self.d_cat), None,
                                            ptr(self.d_kl) if semi else None, self.gtab["enc"],
                                            ptr(self.scratch["encoder"]), st), "encoder backward")

    def _conv_head_call(self, m, what, hst, call, stats_half):
        """One pass of a Conv1d head.  Its BatchNorm1d normalises over the BATCH, the one place where the path is not
        per-sample: on several ranks the pass runs in two phases around an all-reduce of the 2 * C double-precision batch sums
        (SyncBN), so that N ranks x B/N samples normalise exactly like one rank x B.  The dropout masks come from Philox with the
        device-side step counter as offset (a captured graph replays fresh masks every step)."""
        hs = self.head_state[m]
        bn = hs["bn"]

        def state(phase):
            return IinsHeadState(1, None, None, hs["seed"], 0, bn.running_mean.data_ptr(), bn.running_var.data_ptr(),
                                 bn.num_batches_tracked.data_ptr(), hs["stats"].data_ptr(), phase, float(self.world),
                                 self.rank * self.B, self.steps.data_ptr())

        if self.world == 1:
            self.lib.check(call(state(0)), f"{m} (Conv1d) {what}")
            return
        import torch.distributed as dist
        self.lib.check(call(state(1)), f"{m} (Conv1d) {what}, phase 1")
        n = 2 * hs["C"]
        dist.all_reduce(hs["stats"][stats_half * hs["C"]:stats_half * hs["C"] + n], op=dist.ReduceOp.SUM, group=self.pg)
        self.lib.check(call(state(2)), f"{m} (Conv1d) {what}, phase 2")

    # ---- data-parallel gradient exchange ---------------------------------------------------------------
    # The flat gradient buffer is laid out [Enc | Dec | Res | Cls] (module order); the backward produces it back to front:
    # Dec, Res and Cls gradients are complete before the encoder backward starts.  Bucket 1 = everything behind the encoder
    # (only the part that has gradients this step) is summed over the ranks on a communication stream WHILE the encoder
    # backward runs; bucket 0 = the encoder follows on the same stream; the main stream joins before Adam, which applies
    # the 1/world factor (SUM all-reduce + grad_scale = the mean of the per-rank gradients).
    def _bucket_bounds(self, supervised: bool):
        enc_end = self.spans["enc"][1]
        end = self.flat.total if supervised else self.groups[0][1]
        return enc_end, end

    def _allreduce_late_buckets(self, supervised: bool):
        if self.world == 1 or not self.overlap_allreduce:
            return
        import torch.distributed as dist
        enc_end, end = self._bucket_bounds(supervised)
        if end <= enc_end:
            return
        main = torch.cuda.current_stream()
        self.comm_stream.wait_stream(main)
        self._join_weight_gradients(self.comm_stream)          # the Dec / Res / Cls weight gradients may still be in flight
        with torch.cuda.stream(self.comm_stream):
            dist.all_reduce(self.flat.grad[enc_end:end], op=dist.ReduceOp.SUM, group=self.pg)

    def _allreduce(self, supervised: bool):
        if self.world == 1:
            return
        import torch.distributed as dist
        enc_end, end = self._bucket_bounds(supervised)
        if not self.overlap_allreduce:
            dist.all_reduce(self.flat.grad[:end], op=dist.ReduceOp.SUM, group=self.pg)
            return
        main = torch.cuda.current_stream()
        self.comm_stream.wait_stream(main)
        with torch.cuda.stream(self.comm_stream):
            dist.all_reduce(self.flat.grad[:enc_end], op=dist.ReduceOp.SUM, group=self.pg)
        main.wait_stream(self.comm_stream)

    def _adam(self, supervised: bool):
        act = (C.c_int32 * 3)(1, int(supervised), int(supervised))
        fl = self.flat
        self.lib.check(self.lib.iins_adam_step(ptr(fl.flat), ptr(fl.grad), ptr(fl.exp_avg), ptr(fl.exp_avg_sq), self._gb,
                                               self._ge, act, 3, ptr(self.steps), ptr(self.lr), self.betas[0], self.betas[1],
                                               self.eps, 1.0 / self.world, 1, _stream()), "adam")

    def _join_weight_gradients(self, waiter=None):
        """Make ``waiter`` (default: the current stream) wait for the library's weight-gradient streams of the streams this
        engine launched backward passes on (deferred joins, include/iins_b200.h)."""
        main = torch.cuda.current_stream()
        w = C.c_void_p((waiter or main).cuda_stream)
        for st in (main, self.head_stream):
            self.lib.check(self.lib.iins_join_helpers(C.c_void_p(st.cuda_stream), w, int(waiter is not None)), "join_helpers")

    def _step_body(self, supervised: bool, update: bool = True):
        self._forward(supervised)
        self._loss(supervised)
        # the decoder's / heads' weight gradients keep running next to the encoder backward: one join before they are consumed
        self.lib.check(self.lib.iins_set_deferred_join(int(self.concurrent)), "set_deferred_join")
        try:
            self._backward(supervised)
        finally:
            self.lib.check(self.lib.iins_set_deferred_join(0), "set_deferred_join")
        self._join_weight_gradients()
        self._allreduce(supervised)
        if update:
            self._adam(supervised)

    # ------------------------------------------------------------------------------------------------
    def load_batch(self, cir, err, label):
        """Stage one batch (host or device tensors) into the static input buffers (async on the stream)."""
        self.cir.copy_(cir.view(self.B, self.L), non_blocking=True)
        self.err.copy_(err.view(self.B, 1), non_blocking=True)
        self.label.copy_(label.view(self.B, 1).to(torch.float32) if label.dtype != torch.float32 else label.view(self.B, 1),
                         non_blocking=True)

    def prefetch(self, cir, err, label):
        """Input pipeline: start the host->device copy of the NEXT batch on a copy stream into
        staging buffers while the current step computes; ``step(prefetched=True)`` then only moves 2.6 MB device to
        device.  The reference does a synchronous pageable ``.cuda()`` per tensor per step (train_semi.py:174-180)."""
        if not hasattr(self, "_stage"):
            self._stage = (torch.empty_like(self.cir), torch.empty_like(self.err), torch.empty_like(self.label))
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._copy_done = torch.cuda.Event()
            self._stage_free = torch.cuda.Event()
            self._stage_free.record(torch.cuda.current_stream())
        self._copy_stream.wait_event(self._stage_free)          # the previous consumer has read the staging buffers
        with torch.cuda.stream(self._copy_stream):
            self._stage[0].copy_(cir.view(self.B, self.L), non_blocking=True)
            self._stage[1].copy_(err.view(self.B, 1), non_blocking=True)
            lab = label.view(self.B, 1)
            self._stage[2].copy_(lab if lab.dtype == torch.float32 else lab.to(torch.float32), non_blocking=True)
            self._copy_done.record(self._copy_stream)

    def _consume_prefetch(self):
        cur = torch.cuda.current_stream()
        cur.wait_event(self._copy_done)
        self.cir.copy_(self._stage[0], non_blocking=True)
        self.err.copy_(self._stage[1], non_blocking=True)
        self.label.copy_(self._stage[2], non_blocking=True)
        self._stage_free.record(cur)

    def step(self, cir=None, err=None, label=None, supervised=True, update=True, prefetched=False):
        """One optimisation step.  Returns the device tensor ``out`` (8 floats, see iins_b200.h) -- nothing
        synchronises; read ``loss_terms()`` when a host value is needed.  ``prefetched=True`` consumes the batch
        staged by ``prefetch()`` instead of copying ``cir / err / label`` now."""
        if self.mode == "supervised":
            supervised = True
        if prefetched:
            self._consume_prefetch()
        elif cir is not None:
            self.load_batch(cir, err, label)
        key = (bool(supervised), bool(update))
        if getattr(self.flat, "grads_dirty", False):              # the previous step kept its gradients (update=False)
            self.flat.grad.zero_()
            self.flat.grads_dirty = False
        if self.use_graph:
            g = self._graphs.get(key)
            if g is None:
                # warm up once eagerly on a side stream (also validates every call), then capture
                state = (self.flat.flat.clone(), self.flat.exp_avg.clone(), self.flat.exp_avg_sq.clone(), self.steps.clone())
                s = _warmup_stream(self.device)            # one per device: the library keeps helper streams per caller stream
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    self._step_body(*key)
                torch.cuda.current_stream().wait_stream(s)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=_capture_stream(self.device)):
                    self._step_body(*key)
                # the warm-up and the capture pass must not count as optimisation steps
                self.flat.flat.copy_(state[0]); self.flat.exp_avg.copy_(state[1]); self.flat.exp_avg_sq.copy_(state[2])
                self.steps.copy_(state[3])
                self.flat.grad.zero_()
                self._graphs[key] = g
            g.replay()
        else:
            self._step_body(*key)
        self.flat.grads_dirty = not update
        self.n_steps += 1
        return self.out

    def close(self):
        """Drop the captured graphs (they hold the NCCL communicator busy: call before destroy_process_group)."""
        torch.cuda.synchronize(self.device)
        self._graphs.clear()

    def loss_terms(self):
        """Host copy of the loss terms of the last step (this synchronises)."""
        o = self.out.tolist()
        if o[6] != 0.0:
            raise ValueError(f"{int(o[6])} labels outside [0, {self.NC}) after label_offset={self.label_offset} "
                             "(train_semi.py:217-222: labels are 1-based for every dataset_env except 'room_full')")
        kl = float(self.kl) if self.mode == "semi" else 0.0
        lam_res = LAMBDA_RES if self.mode == "semi" else 1.0
        d = dict(loss_ae=LAMBDA_AE * o[0], loss_range=LAMBDA_RANGE * kl, loss_res=lam_res * o[1], loss_env=LAMBDA_ENV * o[2],
                 rmse=float(np.sqrt(max(o[4], 0.0))), mae=o[1], accuracy=o[5] / self.B)
        d["loss"] = o[3] + (LAMBDA_RANGE * kl if self.mode == "semi" else 0.0)
        return d

    # ---- optimizer state (the reference saves only the four module state_dicts, so a resumed
    # run restarts Adam from zero moments -- train_semi.py:281-286; these two methods make the resume exact) -------------
    def optimizer_state_dict(self):
        """``torch.optim.Adam(itertools.chain(Enc, Dec, Res, Cls parameters)).state_dict()`` layout (train_semi.py:118-122):
        ``state[i] = {step, exp_avg, exp_avg_sq}`` for parameter i in chain order, ``param_groups`` with lr / betas / eps.
        Parameters that never received a gradient (restorer.linear_layer2; Res / Cls before the first supervised batch)
        have no entry, exactly like torch's lazily created state.  Loads into a stock torch.optim.Adam."""
        fl = self.flat
        steps = self.steps.tolist()
        state, o = {}, 0
        for i, p in enumerate(fl.params):
            n = p.numel()
            gi = next((g for g, (b, e) in enumerate(self.groups) if b <= o < e), None)
            if gi is not None and steps[gi] > 0:
                state[i] = {"step": torch.tensor(float(steps[gi])), "exp_avg": fl.exp_avg[o:o + n].view(p.shape).clone(),
                            "exp_avg_sq": fl.exp_avg_sq[o:o + n].view(p.shape).clone()}
            o += n
        group = {"lr": float(self.lr.item()), "betas": tuple(self.betas), "eps": self.eps, "weight_decay": 0, "amsgrad": False,
                 "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "decoupled_weight_decay": False, "params": list(range(len(fl.params)))}
        return {"state": state, "param_groups": [group]}

    def load_optimizer_state_dict(self, sd):
        """Inverse of ``optimizer_state_dict`` (also accepts the state_dict of a torch.optim.Adam built over the same
        parameter chain).  Step counters are per Adam group here: the parameters of a group must agree."""
        fl = self.flat
        fl.exp_avg.zero_(); fl.exp_avg_sq.zero_()
        steps = [0] * len(self.groups)
        o = 0
        for i, p in enumerate(fl.params):
            n = p.numel()
            st = sd["state"].get(i)
            if st is not None:
                gi = next((g for g, (b, e) in enumerate(self.groups) if b <= o < e), None)
                if gi is None:
                    raise ValueError(f"parameter {i} has optimizer state but belongs to no Adam group of this engine")
                step = int(st["step"])
                if steps[gi] not in (0, step):
                    raise ValueError("parameters of one Adam group carry different step counts")
                steps[gi] = step
                fl.exp_avg[o:o + n].copy_(st["exp_avg"].reshape(-1))
                fl.exp_avg_sq[o:o + n].copy_(st["exp_avg_sq"].reshape(-1))
            o += n
        self.steps.zero_()
        self.steps[:len(steps)] = torch.tensor(steps, dtype=torch.int32)
        g = sd["param_groups"][0]
        self.set_lr(g["lr"])

    def named_grads(self):
        """{module-prefixed parameter name: gradient view} of the last step (for tests / inspection)."""
        out = {}
        i = 0
        mods = [("enc", self.Enc)] + ([("dec", self.Dec)] if self.mode == "semi" else []) + [("res", self.Res), ("cls", self.Cls)]
        for pre, m in mods:
            for n, _ in m.named_parameters():
                out[f"{pre}.{n}"] = self.flat.grad_views[i]
                i += 1
        return out


class InferenceEngine:
    """test.py:66-85: Encoder -> (Classifier, Restorer), no grad, fused metrics.  Static buffers + CUDA graph."""

    def __init__(self, Enc, Res, Cls, batch_size, cir_len=157, use_graph=True, device=None, label_offset=0):
        self.lib = get_lib()
        self.label_offset = int(label_offset)
        self.device = torch.device(device if device is not None else torch.cuda.current_device())
        o = Enc.opts
        self.B, self.L, self.NC = int(batch_size), int(cir_len), Cls.num_classes
        self.cfg = IinsConfig(self.B, self.L, o["dim"], o["n_residual"], o["n_downsample"], o["env_dim"], o["range_dim"],
                              self.NC, Cls.filters)
        self.lib.check(self.lib.iins_validate_config(self.cfg), "engine config")
        dev, B = self.device, self.B
        f = lambda *shape: torch.zeros(*shape, dtype=torch.float32, device=dev)
        self.cir, self.err, self.label = f(B, self.L), f(B, 1), f(B, 1)
        self.rc, self.cat, self.kl = f(B, o["range_dim"], 128 >> o["n_downsample"]), f(B, o["env_dim"]), f(1)
        self.err_est, self.logits, self.out = f(B, 1), f(B, self.NC), f(8)
        self.pred = torch.zeros(B, dtype=torch.int32, device=dev)
        self.conv_head = {"restorer": getattr(Res, "net_type", "Linear") == "Conv1d",
                          "classifier": getattr(Cls, "net_type", "Linear") == "Conv1d"}
        q = lambda m: f"iins_{m}_conv_ws_floats" if self.conv_head.get(m) else f"iins_{m}_ws_floats"
        self.ws = {m: f(int(getattr(self.lib, q(m))(self.cfg)) + 16) for m in ("encoder", "restorer", "classifier")}
        self.bn = {m: getattr(getattr(mod, m).conv_blocks, "6") for m, mod in (("restorer", Res), ("classifier", Cls)) if self.conv_head[m]}
        self.ptab = {n: ptr_array(list(m.parameters())) for n, m in (("enc", Enc), ("res", Res), ("cls", Cls))}
        self._keep = (Enc, Res, Cls)
        self.use_graph, self._graph = use_graph, None

    def _body(self, with_metrics: bool):
        lib, cfg, st = self.lib, self.cfg, _stream()
        lib.check(lib.iins_encoder_forward(cfg, self.ptab["enc"], ptr(self.cir), None, 0, 0, ptr(self.rc), ptr(self.cat), None,
                                           ptr(self.kl), ptr(self.ws["encoder"]), st), "encoder forward")
        for m, tab, src, dst in (("restorer", "res", self.rc, self.err_est), ("classifier", "cls", self.cat, self.logits)):
            if self.conv_head[m]:                       # eval mode (test.py:43-47): no dropout, BatchNorm on the running statistics
                bn = self.bn[m]
                hs = IinsHeadState(0, None, None, 0, 0, bn.running_mean.data_ptr(), bn.running_var.data_ptr(), None, None, 0, 1.0, 0, None)
                lib.check(getattr(lib, f"iins_{m}_conv_forward")(cfg, self.ptab[tab], ptr(src), ptr(dst), ptr(self.ws[m]), C.byref(hs), st),
                          f"{m} (Conv1d) forward")
            else:
                lib.check(getattr(lib, f"iins_{m}_forward")(cfg, self.ptab[tab], ptr(src), ptr(dst), ptr(self.ws[m]), st), f"{m} forward")
        lib.check(lib.iins_loss_forward_backward(self.B, self.L, self.NC, None, None, ptr(self.err), ptr(self.err_est),
                                                 ptr(self.logits), ptr(self.label), None, self.label_offset, 1.0, 1.0, 1.0,
                                                 ptr(self.out), None, None, None, ptr(self.pred), st), "metrics")

    def close(self):
        """Drop the captured graph (call before destroy_process_group under torchrun)."""
        torch.cuda.synchronize(self.device)
        self._graph = None

    def run(self, cir=None, err=None, label=None):
        """Returns (err_est (B,1), pred (B,) int32, out[8]) as device tensors (static buffers, overwritten next call)."""
        if cir is not None:
            self.cir.copy_(cir.view(self.B, self.L), non_blocking=True)
        if err is not None:
            self.err.copy_(err.view(self.B, 1), non_blocking=True)
            self.label.copy_(label.view(self.B, 1).float(), non_blocking=True)
        if self.use_graph:
            if self._graph is None:
                s = _warmup_stream(self.device)
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    self._body(True)
                torch.cuda.current_stream().wait_stream(s)
                self._graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self._graph):
                    self._body(True)
            self._graph.replay()
        else:
            self._body(True)
        return self.err_est, self.pred, self.out
