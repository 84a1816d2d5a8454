"""Input pipeline: zenodo pickle -> split -> StandardScaler -> pinned-memory batch ring -> async H2D.

Restates what the reference's loaders are meant to do (its own files do not parse: data_tools.py has a SyntaxError at :47,
dataset.py dies importing it) with the same names and return shapes:

  * ``load_pkl_data(filepath, option)``         data_tools.py:114-337 -- the pandas pickle of the Deep UWB dataset
    (README_diverse.md: columns CIR (157 floats), Error (m), Room (int 0-4), Obstacles (10-character one-hot string))
    -> (cir (n,157), err (n,1), label (n,1), room (n,1)), shuffled;
  * ``err_mitigation_dataset(root, ...)``        dataset.py:15-89 -- 'full' split by ``split_factor`` or the 'paper' split
    (medium room = label 2 held out), optional StandardScaler fitted on the training CIRs (:73-76);
  * ``UWBDataset``                               dataset.py:92-136 -- ``{"CIR", "Err", "Label"}`` float32 items;
  * ``PinnedBatchRing``                          replaces ``DataLoader(UWBDataset(...), shuffle=True, num_workers=8)``
    (train_semi.py:142-154): a background thread gathers each shuffled batch straight into a ring of PINNED host buffers, so
    that ``SemiTrainEngine.prefetch`` can start the host-to-device copy of batch i+1 while step i computes (the reference
    pickles every batch through worker pipes and then does a synchronous pageable ``.cuda()`` per tensor, :174-180).
"""
import queue
import threading

import numpy as np
import torch


class StandardScaler:
    """sklearn.preprocessing.StandardScaler semantics (dataset.py:73-76): per-feature mean and POPULATION standard deviation
    of the training set; a zero-variance feature is left unscaled."""

    def fit(self, x):
        x = np.asarray(x, dtype=np.float64)
        self.mean_ = x.mean(axis=0)
        self.var_ = x.var(axis=0)
        self.scale_ = np.sqrt(self.var_)
        self.scale_[self.scale_ == 0.0] = 1.0
        return self

    def transform(self, x):
        return (np.asarray(x, dtype=np.float64) - self.mean_) / self.scale_

    def fit_transform(self, x):
        return self.fit(x).transform(x)


# obstacle one-hot strings of the four materials of the 'obstacle_full' option (data_tools.py:250-291)
_OBSTACLE_FULL = ["0000000001", "0000000100", "0010000000", "0000000010"]


def load_pkl_data(filepath, option=None, rng=None):
    """data_tools.py:114-337 for the options whose branch is well defined there: 'room_full' (labels = Room 0..4, :161-168) and
    'obstacle_full' (labels 0..3 for the four obstacle materials, :248-300).  Returns (cir, err, label, room), jointly shuffled
    (np.random.shuffle in the reference; pass ``rng`` for a reproducible order)."""
    import pandas as pd
    rng = rng if rng is not None else np.random
    data = pd.read_pickle(filepath)
    if option in (None, "room_full"):
        cir = np.vstack(data["CIR"].to_numpy())
        err = np.asarray(data["Error"], dtype=np.float64).reshape(-1, 1)
        room = np.asarray(data["Room"], dtype=np.float64).reshape(-1, 1)
        label = room.copy()
    elif option == "obstacle_full":
        parts = []
        for k, code in enumerate(_OBSTACLE_FULL):
            ds = data.loc[data["Obstacles"] == code]
            n = len(ds)
            parts.append((np.vstack(ds["CIR"].to_numpy()) if n else np.zeros((0, 157)),
                          np.asarray(ds["Error"], dtype=np.float64).reshape(-1, 1), np.full((n, 1), float(k)),
                          np.asarray(ds["Room"], dtype=np.float64).reshape(-1, 1)))
        cir, err, label, room = (np.vstack([p[i] for p in parts]) for i in range(4))
    else:
        raise NotImplementedError(f"load_pkl_data: option {option!r} (the reference's branch for it is not well defined)")
    perm = rng.permutation(len(cir))
    return cir[perm], err[perm], label[perm], room[perm]


def err_mitigation_dataset(root, dataset_name="zenodo", dataset_env=None, split_factor=0.8, scaling=False, mode="paper",
                           feature_flag=False, data=None, rng=None):
    """dataset.py:15-89.  ``data`` = (cir, err, label) arrays may be passed instead of a pickle path.  Returns
    (train, test, None, None) with train / test = (cir (n,L), err (n,1), label (n,1))."""
    if data is None:
        if dataset_name != "zenodo":
            raise NotImplementedError("only the zenodo dataset is wired (the reference marks ewine 'not used', dataset.py:22)")
        cir, err, label, _ = load_pkl_data(root, dataset_env or "room_full", rng)
    else:
        cir, err, label = (np.asarray(a, dtype=np.float64) for a in data)
    err, label = err.reshape(len(err), 1), label.reshape(len(label), 1)
    if mode == "full":
        k = int(len(err) * split_factor)
        train, test = (cir[:k], err[:k], label[:k]), (cir[k:], err[k:], label[k:])
    elif mode == "paper":                                     # medium room (label 2) is the test set (dataset.py:37-56)
        m = label[:, -1] == 2
        train, test = (cir[~m], err[~m], label[~m]), (cir[m], err[m], label[m])
    else:
        raise ValueError(f"unknown mode {mode!r}")
    if scaling:
        sc = StandardScaler()
        train = (sc.fit_transform(train[0]), train[1], train[2])
        test = (sc.transform(test[0]), test[1], test[2])
    return train, test, None, None


class UWBDataset(torch.utils.data.Dataset):
    """dataset.py:92-136."""

    def __init__(self, data):
        self.data = data
        self.cir, self.err, self.label = (np.ascontiguousarray(a, dtype=np.float32) for a in data)
        self.err, self.label = self.err.reshape(len(self.err), 1), self.label.reshape(len(self.label), 1)

    def __getitem__(self, index):
        i = index % len(self.cir)
        return {"CIR": torch.from_numpy(self.cir[i]), "Err": torch.from_numpy(self.err[i]), "Label": torch.from_numpy(self.label[i])}

    def __len__(self):
        return len(self.cir)


class PinnedBatchRing:
    """``DataLoader(dataset, batch_size, shuffle=True)`` for a UWBDataset, without worker processes: one producer thread
    gathers every batch of the epoch's permutation into the next free slot of a ring of pinned host buffers
    (numpy fancy-index -> pinned tensor, no pickling, no per-sample tensors) and hands the slot to the consumer.
    The consumer copies out of a slot ASYNCHRONOUSLY (``SemiTrainEngine.prefetch`` / ``load_batch``), so a slot is recycled
    only behind a CUDA event: when batch k is requested, the steps of all batches <= k-2 have been enqueued on the current
    stream (the training loop looks one batch ahead), an event recorded there covers their host-to-device copies, and the
    slot of batch k-2 goes back to the producer once that event has completed.  This also bounds how far the host may run
    ahead of the device (about two steps).  The ring holds ``depth`` >= 4 slots."""

    def __init__(self, dataset, batch_size, shuffle=True, drop_last=False, depth=4, seed=None, pin=None):
        self.dataset, self.batch_size, self.shuffle, self.drop_last = dataset, int(batch_size), shuffle, drop_last
        self.depth = max(4, int(depth))
        self.cuda_fence = bool(pin) if pin is not None else torch.cuda.is_available()
        self.rng = np.random.RandomState(seed)
        pin = torch.cuda.is_available() if pin is None else pin
        L = dataset.cir.shape[1]
        mk = lambda *s: torch.empty(*s, dtype=torch.float32).pin_memory() if pin else torch.empty(*s, dtype=torch.float32)
        self.slots = [(mk(self.batch_size, L), mk(self.batch_size, 1), mk(self.batch_size, 1)) for _ in range(self.depth)]

    def __len__(self):
        n = len(self.dataset)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        n = len(self.dataset)
        order = self.rng.permutation(n) if self.shuffle else np.arange(n)
        nb = len(self)
        free, full = queue.Queue(), queue.Queue(maxsize=self.depth)
        if self.cuda_fence and torch.cuda.is_available():
            torch.cuda.synchronize()      # epoch boundary: copies out of the previous epoch's last slots may still be in flight
        for s in range(self.depth):
            free.put(s)
        ds = self.dataset

        def produce():
            for b in range(nb):
                idx = np.sort(order[b * self.batch_size:(b + 1) * self.batch_size]) if not self.shuffle else order[b * self.batch_size:(b + 1) * self.batch_size]
                s = free.get()
                if s is None:
                    return
                k = len(idx)
                cir, err, lab = self.slots[s]
                np.take(ds.cir, idx, axis=0, out=cir.numpy()[:k])
                np.take(ds.err, idx, axis=0, out=err.numpy()[:k])
                np.take(ds.label, idx, axis=0, out=lab.numpy()[:k])
                full.put((s, k))
            full.put(None)

        t = threading.Thread(target=produce, daemon=True)
        t.start()
        held = []
        try:
            while True:
                item = full.get()
                if item is None:
                    break
                s, k = item
                held.append(s)
                if len(held) > 2:                              # the slot two batches back: fence, then hand it to the producer
                    if self.cuda_fence and torch.cuda.is_available():
                        ev = torch.cuda.Event()
                        ev.record(torch.cuda.current_stream())
                        ev.synchronize()
                    free.put(held.pop(0))
                cir, err, lab = self.slots[s]
                yield {"CIR": cir[:k], "Err": err[:k], "Label": lab[:k]}
        finally:
            free.put(None)
