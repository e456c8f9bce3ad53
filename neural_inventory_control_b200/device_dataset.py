"""Device-resident dataset + index-gather batch assembly.

The reference feeds `Trainer.do_one_epoch` from a torch DataLoader over `MyDataset`: every batch is built by
calling `__getitem__` once per sample and `default_collate`-ing B dicts on the host, then copied H2D
(`data_handling.py:385-395`, `main_run.py:101-104`, `trainer.py:155-156`) - 55 % of a CPU epoch in the reference and
~100x the fused rollout's own time here. `DeviceBatches` keeps the DataLoader's semantics (batch size, fresh random
permutation per epoch when shuffle=True, last partial batch kept unless drop_last) but holds the tensors in HBM and
assembles each batch with one gather kernel per tensor (`hdpo_gather_rows`).
"""
import torch
from torch.utils.data import RandomSampler, SequentialSampler

from . import _capi as K
from . import _lib
from .engine import current_stream_ptr


def eligible(loader):
    ds = getattr(loader, "dataset", None)
    data = getattr(ds, "data", None)
    if not isinstance(data, dict) or not data or not all(torch.is_tensor(v) for v in data.values()):
        return False
    if loader.batch_size is None or getattr(loader, "batch_sampler", None) is None:
        return False
    return isinstance(loader.sampler, (RandomSampler, SequentialSampler))


class DeviceBatches:
    def __init__(self, loader, device):
        self.loader = loader
        self.dataset = loader.dataset
        self.device = torch.device(device)
        self.batch_size = loader.batch_size
        self.drop_last = loader.drop_last
        self.shuffle = isinstance(loader.sampler, RandomSampler)
        self.n = len(loader.dataset)
        self.lib = _lib.load()
        # one H2D copy per tensor for the lifetime of the loader (expanded views are materialised); dtypes are kept
        # (float32 rows go through the gather kernel, anything else through torch.index_select)
        self.data = {k: v.to(self.device).contiguous() for k, v in loader.dataset.data.items()}

    def __len__(self):
        full, rem = divmod(self.n, self.batch_size)
        return full + (1 if rem and not self.drop_last else 0)

    def gather(self, index):
        out = {}
        stream = current_stream_ptr(self.device)
        for k, src in self.data.items():
            if src.dtype != torch.float32:
                out[k] = src.index_select(0, index)
                continue
            row = src[0].numel() if src.dim() > 1 else 1
            dst = torch.empty((index.numel(),) + tuple(src.shape[1:]), dtype=torch.float32, device=self.device)
            rc = self.lib.hdpo_gather_rows(dst.data_ptr(), src.data_ptr(), index.data_ptr(), index.numel(), row, stream)
            K.check(self.lib, rc, "hdpo_gather_rows")
            out[k] = dst
        return out

    def _permutation(self):
        """The permutation DataLoader + RandomSampler would draw, consuming the SAME host RNG draws in the same order
        (torch/utils/data/dataloader.py: `_BaseDataLoaderIter.__init__` draws base_seed from loader.generator;
        sampler.py: RandomSampler seeds a fresh generator from the global RNG when it has none), so that batch order
        and every later draw from the global generator (e.g. LazyLinear initialisation) match the reference run.
        Under torch.distributed rank 0's permutation is broadcast: all ranks must cut the same batches."""
        torch.empty((), dtype=torch.int64).random_(generator=getattr(self.loader, "generator", None))  # base_seed
        gen = getattr(self.loader.sampler, "generator", None)
        if gen is None:
            seed = int(torch.empty((), dtype=torch.int64).random_().item())
            gen = torch.Generator()
            gen.manual_seed(seed)
        perm = torch.randperm(self.n, generator=gen)
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            buf = perm.to(self.device) if dist.get_backend() == "nccl" else perm
            dist.broadcast(buf, src=0)
            perm = buf.cpu()
        return perm

    def __iter__(self):
        if not self.shuffle:
            # a DataLoader iterator draws its base_seed even when nothing is shuffled: keep the global RNG in step
            torch.empty((), dtype=torch.int64).random_(generator=getattr(self.loader, "generator", None))
            for a in range(0, self.n, self.batch_size):
                b = min(a + self.batch_size, self.n)
                if b - a < self.batch_size and self.drop_last:
                    return
                yield {k: v[a:b] for k, v in self.data.items()}  # contiguous views, no copy
            return
        perm = self._permutation().to(self.device)
        for a in range(0, self.n, self.batch_size):
            b = min(a + self.batch_size, self.n)
            if b - a < self.batch_size and self.drop_last:
                return
            yield self.gather(perm[a:b].contiguous())
