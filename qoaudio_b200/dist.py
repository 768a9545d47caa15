"""torch.distributed plumbing for the one-process-per-GPU launch (NCCL on the GPU box, gloo in CPU tests).

Nothing here touches the data path: it provides the barrier and the max / sum reductions bench.py needs to report a
whole-job number (units processed by all ranks / slowest rank's time)."""
import os


class Group:
    def __init__(self, backend=None, device=None):
        import torch
        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.dist = None
        self.device = device if device is not None else torch.device("cpu")
        if self.world > 1:
            import torch.distributed as dist
            if not dist.is_initialized():
                backend = backend or ("nccl" if self.device.type == "cuda" else "gloo")
                kw = {"device_id": self.device} if backend == "nccl" else {}
                dist.init_process_group(backend, **kw)
            self.dist = dist

    def barrier(self):
        if self.device.type == "cuda":
            self.torch.cuda.synchronize(self.device)
        if self.dist is not None:
            self.dist.barrier()
        if self.device.type == "cuda":
            self.torch.cuda.synchronize(self.device)

    def _reduce(self, x, op):
        if self.dist is None:
            return float(x)
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device=self.device)
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, x):
        return self._reduce(x, self.dist.ReduceOp.MAX if self.dist else None)

    def sum(self, x):
        return self._reduce(x, self.dist.ReduceOp.SUM if self.dist else None)

    def throughput(self, units_local, seconds_local):
        """Whole-job throughput: units of all ranks / the slowest rank's time."""
        return self.sum(units_local) / self.max(seconds_local)

    def close(self):
        if self.dist is not None and self.dist.is_initialized():
            self.dist.destroy_process_group()
