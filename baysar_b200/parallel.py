"""Data-parallel evaluation of a theta batch over the GPUs of one box.

Every theta is an independent unit and the model constants (a few MB) are replicated per rank, so the batch is cut
into contiguous row blocks, one per rank, with NO collective on the data path.  The only exchange is the one an
ensemble sampler needs: an all-gather of the per-walker log-posteriors (8 bytes per walker) so that every rank sees
all walkers for the accept step.  One process per GPU; `torch.distributed` (NCCL over NVLink for CUDA tensors, gloo
in the CPU tests) is plumbing only.
"""
import torch
import torch.distributed as dist


def shard_bounds(n_total, world, rank):
    """Contiguous, balanced row block [lo, hi) of rank `rank`: the first n_total % world ranks get one extra row."""
    base, extra = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class ShardedEvaluator:
    """Evaluate ``eval_fn`` on this rank's rows of a replicated theta batch and all-gather the results.

    ``eval_fn(theta_rows) -> 1-D tensor`` is ``BaysarPosterior.logp_batch`` in production; the CPU tests pass a
    stand-in so that the sharding / padding / gather logic is exercised without a GPU.
    """

    def __init__(self, eval_fn, group=None):
        self.eval_fn = eval_fn
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._side = None  # side stream + receive buffers of the overlapped all-gather (evaluate_local(overlap=True))

    def local_rows(self, n_total):
        return shard_bounds(n_total, self.world, self.rank)

    def __call__(self, theta):
        """theta: (N, n_params) tensor, identical on every rank -> (N,) tensor of results on every rank."""
        n = theta.shape[0]
        lo, hi = self.local_rows(n)
        local = self.eval_fn(theta[lo:hi]) if hi > lo else theta.new_empty(0)
        if self.world == 1:
            return local
        width = -(-n // self.world)  # rows per rank after padding to equal size
        send = local.new_full((width,), float("nan"))
        send[: hi - lo] = local
        recv = local.new_empty(self.world * width)
        dist.all_gather_into_tensor(recv, send, group=self.group)
        parts = []
        for r in range(self.world):
            rlo, rhi = shard_bounds(n, self.world, r)
            parts.append(recv[r * width: r * width + (rhi - rlo)])
        return torch.cat(parts)

    def evaluate_local(self, theta_rows, overlap=False):
        """Weak-scaling form used by bench.py: every rank brings its own rows; returns the gathered vector.

        ``overlap=True``: the all-gather runs on a side stream behind an event, so that it overlaps with whatever the
        caller enqueues next on the current stream (the next batch of an independent-walker ensemble does not depend on
        it); the returned tensor is valid once :meth:`wait` has been called (or the device synchronised).  Two receive
        buffers alternate, so at most one gather may be outstanding when the next one is issued — `evaluate_local`
        makes the side stream wait for that itself.
        """
        local = self.eval_fn(theta_rows)
        if self.world == 1:
            return local
        if not overlap:
            recv = local.new_empty(self.world * local.shape[0])
            dist.all_gather_into_tensor(recv, local.contiguous(), group=self.group)
            return recv
        if self._side is None:
            self._side = torch.cuda.Stream(device=local.device)
            self._recv = [None, None]
            self._turn = 0
        n = self.world * local.shape[0]
        k = self._turn
        if self._recv[k] is None or self._recv[k].numel() != n:
            self._recv[k] = local.new_empty(n)
        ready = torch.cuda.Event()
        ready.record()                       # local logP complete on the current stream
        local = local.contiguous()
        local.record_stream(self._side)      # keep the allocator from recycling it before the gather has read it
        with torch.cuda.stream(self._side):
            self._side.wait_event(ready)
            dist.all_gather_into_tensor(self._recv[k], local, group=self.group)
        self._turn ^= 1
        return self._recv[k]

    def wait(self):
        """Make the current stream wait for the outstanding overlapped all-gathers."""
        if self._side is not None:
            torch.cuda.current_stream().wait_stream(self._side)


def stretch_move_proposal(walkers, complement, generator=None, a=2.0):
    """Affine-invariant stretch move (Goodman & Weare) for one half of a red/blue ensemble split — the caller of the
    batched posterior for the 1M-walker workload (reference caller: emcee in `tulasa/fitting.py:1275-1304`).

    walkers (S, n), complement (C, n) -> (proposal (S, n), log-Hastings factor (S,)): accept with probability
    min(1, exp(factor + logp_new - logp_old)).  Pure tensor ops: runs on whatever device the walkers live on.
    """
    s, n = walkers.shape
    u = torch.rand(s, generator=generator, device=walkers.device, dtype=walkers.dtype)
    z = ((a - 1.0) * u + 1.0) ** 2 / a
    pick = torch.randint(complement.shape[0], (s,), generator=generator, device=walkers.device)
    partner = complement[pick]
    proposal = partner + z[:, None] * (walkers - partner)
    return proposal, (n - 1) * torch.log(z)
