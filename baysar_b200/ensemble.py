"""On-device affine-invariant ensemble sampler (stretch move, red/blue split) around the batched posterior.

The reference drives emcee's ``EnsembleSampler`` from the host (`tulasa/fitting.py:1275-1304`): every half-step costs
a host round trip per walker.  Here walkers, proposals and log-posteriors stay in HBM: per half-step one proposal
kernel, the batched posterior (LOS -> chord -> ... -> finalize) and one accept kernel, all on the current stream
(`csrc/ensemble.cuh`; Philox4x32-10 keyed by ``seed`` and counted by GLOBAL walker index, so a run is reproducible from
``seed`` alone and ranks never share a stream).  With `torch.distributed` the walkers are sharded over the ranks and
form ONE ensemble (``global_ensemble=True``, the default under a process group): before each half-step the
complementary half of every rank is all-gathered over NCCL (N/2 x n_params x 8 bytes in total) and the stretch move
draws its partner from all of it, which is what "one ensemble of 2^20 walkers on 8 GPUs" (BASELINE configs[3]) means;
``global_ensemble=False`` keeps one independent ensemble per rank (no data-path collective).  `gather_logp`
all-gathers the per-walker log-posteriors, which is all a convergence monitor needs.
"""
import ctypes

import numpy as np

from . import runtime

_M0, _M1, _W0, _W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85


def philox4x32_10(counter, key):
    """Reference implementation of the kernel's generator (numpy, vectorised over the first counter word)."""
    c = [np.asarray(x, dtype=np.uint64) & 0xFFFFFFFF for x in counter]
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = _M0 * c[0], _M1 * c[2]
        c = [((p1 >> np.uint64(32)) ^ c[1] ^ k0) & 0xFFFFFFFF, p1 & 0xFFFFFFFF,
             ((p0 >> np.uint64(32)) ^ c[3] ^ k1) & 0xFFFFFFFF, p0 & 0xFFFFFFFF]
        k0, k1 = (k0 + _W0) & 0xFFFFFFFF, (k1 + _W1) & 0xFFFFFFFF
    return c


def uniform01(hi, lo):
    bits = ((np.asarray(hi, dtype=np.uint64) << np.uint64(32)) | np.asarray(lo, dtype=np.uint64)) >> np.uint64(12)
    return (bits.astype(np.float64) + 0.5) / 4503599627370496.0


def reference_draws(n_walkers, seed, step, n_complement, a=2.0, walker_offset=0):
    """(z, partner index, ln u_accept) the kernels draw for the walkers with global indices walker_offset ..
    walker_offset + n_walkers - 1 in one half-step."""
    s = np.arange(n_walkers, dtype=np.uint64) + np.uint64(walker_offset)
    key = (seed & 0xFFFFFFFF, seed >> 32)
    zero = np.zeros(n_walkers, dtype=np.uint64)
    hi = 2 * ((step >> 32) & 0xFFFFFFFF)
    r = philox4x32_10((s & 0xFFFFFFFF, s >> np.uint64(32), zero + (step & 0xFFFFFFFF), zero + hi), key)
    t = (a - 1.0) * uniform01(r[0], r[1]) + 1.0
    pick = np.minimum((uniform01(r[2], r[3]) * n_complement).astype(np.int64), n_complement - 1)
    q = philox4x32_10((s & 0xFFFFFFFF, s >> np.uint64(32), zero + (step & 0xFFFFFFFF), zero + hi + 1), key)
    return t * t / a, pick, np.log(uniform01(q[0], q[1]))


def gather_complement(local, world=None):
    """The complementary half-ensemble a half-step draws partners from: this rank's rows followed by nothing else without
    a process group, else the rows of every rank in rank order (one all-gather; works on NCCL and gloo tensors)."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    out = local.new_empty((dist.get_world_size() * local.shape[0],) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, local.contiguous())
    return out


class StretchEnsemble:
    """``StretchEnsemble(posterior, walkers)``: walkers (N, n_params) array or CUDA tensor, N even."""

    def __init__(self, posterior, walkers, seed=0, a=2.0, global_ensemble=None):
        import torch
        import torch.distributed as dist

        self.torch, self.post, self.a, self.seed = torch, posterior, float(a), int(seed)
        grouped = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.rank, self.world = (dist.get_rank(), dist.get_world_size()) if grouped else (0, 1)
        # one ensemble over all ranks (complement all-gathered) unless asked for rank-local ensembles
        self.global_ensemble = grouped if global_ensemble is None else bool(global_ensemble) and grouped
        self.lib = runtime.library()
        self.walkers = posterior._as_device(walkers).clone().contiguous()
        n, self.n_params = self.walkers.shape
        if n % 2 or n < 4:
            raise ValueError("the red/blue split needs an even number (>= 4) of walkers")
        self.half = n // 2
        logp, status = posterior.logp_batch(self.walkers, return_status=True)
        if bool((status != 0).any()) or not bool(torch.isfinite(logp).all()):
            raise ValueError("every starting walker must evaluate cleanly (status 0, finite logP)")
        self.logp = logp.clone()
        self.proposal = torch.empty((self.half, self.n_params), dtype=torch.float64, device=self.walkers.device)
        self.log_factor = torch.empty(self.half, dtype=torch.float64, device=self.walkers.device)
        self.accepted = torch.zeros(1, dtype=torch.int64, device=self.walkers.device)
        self.iteration = 0

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError("baysar_b200: " + self.lib.baysar_last_error().decode())

    def _half_step(self, lo, hi, clo, chi, counter):
        w, c = self.walkers[lo:hi], self.walkers[clo:chi]
        if self.global_ensemble:
            c = gather_complement(c)  # every rank's complementary half: partners come from the whole ensemble
        # global index of this launch's first walker: rank-major within each half (distinct streams on every rank)
        offset = self.rank * self.half + (0 if lo == 0 else self.world * self.half)
        stream = ctypes.c_void_p(self.torch.cuda.current_stream(self.walkers.device).cuda_stream)
        self._check(self.lib.baysar_stretch_propose(w.data_ptr(), c.data_ptr(), self.n_params, hi - lo, c.shape[0], self.a,
                                                    self.seed, counter, offset, self.proposal.data_ptr(),
                                                    self.log_factor.data_ptr(), stream))
        logp_new, status = self.post.logp_batch(self.proposal, return_status=True)
        self._check(self.lib.baysar_stretch_accept(w.data_ptr(), self.logp[lo:hi].data_ptr(), self.proposal.data_ptr(),
                                                   logp_new.data_ptr(), self.log_factor.data_ptr(), status.data_ptr(),
                                                   self.n_params, hi - lo, self.seed, counter, offset,
                                                   self.accepted.data_ptr(), stream))

    def step(self):
        """One ensemble step = red half against blue, then blue half against the updated red."""
        h = self.half
        self._half_step(0, h, h, 2 * h, 2 * self.iteration)
        self._half_step(h, 2 * h, 0, h, 2 * self.iteration + 1)
        self.iteration += 1

    def run(self, n_steps):
        for _ in range(n_steps):
            self.step()
        return self.walkers, self.logp

    @property
    def acceptance_fraction(self):
        return float(self.accepted.item()) / max(1, 2 * self.half * self.iteration)

    def gather_logp(self):
        """Per-walker log-posteriors of every rank (NCCL all-gather); this rank's alone without a process group."""
        import torch.distributed as dist

        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return self.logp
        out = self.logp.new_empty(dist.get_world_size() * self.logp.shape[0])
        dist.all_gather_into_tensor(out, self.logp.contiguous())
        return out
