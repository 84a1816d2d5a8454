"""Batches of the shape ``UWBDataset.__getitem__`` yields (dataset.py:118-133): dicts with "CIR" (B,L) f32,
"Err" (B,1) f32, "Label" (B,1) f32 holding integers.  The zenodo pickle is not shipped with the reference
(the loaders are out of scope and do not parse), so the entry points run on synthetic
tensors of that shape unless the caller passes its own DataLoader."""
import torch


class SyntheticCIR:
    """An iterable of ``n_samples // batch_size`` pinned-memory batches; statistics per SURVEY.md 8(d):
    CIR ~ N(0,1) (StandardScaler'd, dataset.py:73-76), Err = clip(|N(0,0.15)|, 0, 1), Label uniform in [0,NC)."""

    def __init__(self, n_samples, batch_size, cir_len=157, num_classes=5, seed=1234, pin=True, label_base=0):
        g = torch.Generator().manual_seed(seed)
        n = (n_samples // batch_size) * batch_size
        self.batch_size = batch_size
        self.cir = torch.randn(n, cir_len, generator=g)
        self.err = (torch.randn(n, 1, generator=g) * 0.15).abs().clamp_(0, 1)
        # label_base=1: the 1..NC labels of every dataset_env except 'room_full' (train_semi.py:217-222 subtracts 1)
        self.label = (torch.randint(0, num_classes, (n, 1), generator=g) + int(label_base)).float()
        if pin and torch.cuda.is_available():
            self.cir, self.err, self.label = self.cir.pin_memory(), self.err.pin_memory(), self.label.pin_memory()

    def __len__(self):
        return self.cir.shape[0] // self.batch_size

    def __iter__(self):
        b = self.batch_size
        for i in range(len(self)):
            s = slice(i * b, (i + 1) * b)
            yield {"CIR": self.cir[s], "Err": self.err[s], "Label": self.label[s]}
