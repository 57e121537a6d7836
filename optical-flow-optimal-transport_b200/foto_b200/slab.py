"""Time-slab mode: ONE huge FOTO volume split over the ranks of a torch.distributed (NCCL) group.

SURVEY.md section 8(e), second row / BASELINE.json config 5.  Rank g owns the contiguous time planes
[n0_g, n1_g) of every field.  Per outer ALG2 iteration the ranks exchange

  * the boundary planes of mu_rho and q_a (halo of the time difference in K1)        -- isend/irecv
  * the t-slab <-> y-slab transpose of the spectrum, forward and back (K2b: the x and y transforms are
    slab-local, the t transform needs all planes of a pixel)                         -- all_to_all_single
  * the boundary planes of phi (halo of the time difference in K3)                    -- isend/irecv
  * the two sums of the stopping criterion                                            -- all_reduce

and run the library's slab kernels in between (foto_slab_rhs_dev, foto_dct_xy_dev, foto_dct_t_solve_dev,
foto_slab_prox_dev on the torch stream).  The default Poisson back-end is the exact DCT solve: it is the only
solver that needs no per-CG-iteration collective.  Every field value is computed with the same arithmetic
as on one GPU, so the gathered result is bit-identical to `foto_b200.solve(..., backend=POISSON_DCT_EXACT)`
whenever the outer-iteration count agrees (the criterion is summed in a different order).

`poisson="cg_parity"` runs the reference's truncated CG instead (benamou_brenier.py:85 semantics across the ranks,
foto_slab_cg_dev): per CG iteration one boundary plane of r to each neighbour and two one-word all-reduces; the
host enqueues iterations ahead and reads the device's `done` flag every 16 iterations.  Same recurrences as the
one-GPU streaming kernel, dot products summed in another order: 1e-9 agreement unless a CG count flips by one.

The product path needs a GPU per rank and the NCCL back-end.  With a gloo process group (which cannot move CUDA
tensors point to point) the same exchanges are staged through host memory: slower, but it lets two ranks share ONE
GPU, which is how the 2-rank bit-identity test runs on a single-GPU box.  `plan()` (pure Python) is what the CPU
tests cover.
"""
import math

import numpy as np

from .lib import Context as _Context


def split(n, world):
    """Contiguous partition of range(n) over `world` ranks: [(start, stop), ...] (sizes differ by <= 1)."""
    return [((g * n) // world, ((g + 1) * n) // world) for g in range(world)]


def plan(Nt, Ny, world):
    """Slab geometry: time planes and (for the transposed t solve) image rows owned by each rank."""
    if world < 1 or Nt < world or Ny < world:
        raise ValueError(f"cannot split Nt={Nt}, Ny={Ny} over {world} ranks (every rank needs a plane and a row)")
    return {"t": split(Nt, world), "y": split(Ny, world)}


class SlabSolver:
    def __init__(self, Nt, Nx, Ny, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.Nt, self.Nx, self.Ny, self.P = int(Nt), int(Nx), int(Ny), int(Nx) * int(Ny)
        self.staged = dist.is_initialized() and dist.get_backend() != "nccl"      # gloo: exchanges go through host memory
        self.geom = plan(self.Nt, self.Ny, self.world)
        self.n0, self.n1 = self.geom["t"][self.rank]
        self.y0, self.y1 = self.geom["y"][self.rank]
        self.nloc, self.nyl = self.n1 - self.n0, self.y1 - self.y0
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        self.ctx = _Context(self.dev.index)
        self.ctx.set_stream(torch.cuda.current_stream(self.dev).cuda_stream)
        f64 = dict(dtype=torch.float64, device=self.dev)
        L, P = self.nloc, self.P
        self.mu = torch.zeros((3, L + 2, P), **f64)          # planes 0 and L+1 are halos
        self.q = torch.zeros((3, L + 2, P), **f64)
        self.phi = torch.zeros((L + 2, P), **f64)
        self.F = torch.empty((L, P), **f64)
        self.A = torch.empty((L, P), **f64)                   # spectrum, t-slab layout
        self.tmp = torch.empty((L, P), **f64)
        self.B = torch.empty((self.Nt, self.nyl * self.Nx), **f64)   # spectrum, y-slab layout
        self.B2 = torch.empty_like(self.B)
        self.xbuf = torch.empty(L * P, **f64)                 # all-to-all staging: block of rank g = [L][rows of g][Nx]
        self.sums = torch.zeros(2, **f64)
        self.cs = (L + 2) * P
        self._cg = None                                       # buffers of the cg_parity back-end, allocated on first use

    # ------------------------------------------------------------------ exchanges
    def _halo(self, fields):
        """fields: tensors of shape [L+2, P]; fill plane 0 from rank-1's last owned plane and plane L+1 from
        rank+1's first owned plane."""
        if self.world == 1:
            return
        dist, L = self.dist, self.nloc
        ops, back = [], []

        def snd(t, peer):
            ops.append(dist.P2POp(dist.isend, t.cpu() if self.staged else t, peer))

        def rcv(t, peer):
            buf = self.torch.empty(t.shape, dtype=t.dtype) if self.staged else t
            ops.append(dist.P2POp(dist.irecv, buf, peer))
            if self.staged:
                back.append((t, buf))

        for f in fields:
            if self.rank > 0:
                snd(f[1], self.rank - 1); rcv(f[0], self.rank - 1)
            if self.rank < self.world - 1:
                snd(f[L], self.rank + 1); rcv(f[L + 1], self.rank + 1)
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        for t, buf in back:
            t.copy_(buf)

    def _a2a(self, out, inp, out_split, in_split):
        if not self.staged:
            self.dist.all_to_all_single(out, inp, out_split, in_split)
            return
        host = self.torch.empty(out.numel(), dtype=out.dtype)
        self.dist.all_to_all_single(host, inp.cpu(), out_split, in_split)
        out.copy_(host)

    def _to_y_slabs(self, A, B):
        """A: [L, Ny, Nx] (my planes, all rows) -> B: [Nt, nyl, Nx] (all planes, my rows)."""
        torch, dist = self.torch, self.dist
        A3 = A.view(self.nloc, self.Ny, self.Nx)
        if self.world == 1:
            B.view(self.Nt, self.Ny, self.Nx).copy_(A3)
            return
        self.ctx.slab_pack(0, self.nloc, self.Ny, self.Nx, self.world, A.data_ptr(), self.xbuf.data_ptr())   # one gather kernel
        in_split = [self.nloc * (y1 - y0) * self.Nx for (y0, y1) in self.geom["y"]]
        out_split = [(n1 - n0) * self.nyl * self.Nx for (n0, n1) in self.geom["t"]]
        self._a2a(B.view(-1), self.xbuf, out_split, in_split)                 # rank order = plane order

    def _to_t_slabs(self, B, A):
        """B: [Nt, nyl, Nx] -> A: [L, Ny, Nx]."""
        torch, dist = self.torch, self.dist
        if self.world == 1:
            A.view(self.nloc, self.Ny, self.Nx).copy_(B.view(self.Nt, self.Ny, self.Nx))
            return
        in_split = [(n1 - n0) * self.nyl * self.Nx for (n0, n1) in self.geom["t"]]      # contiguous plane ranges of B
        out_split = [self.nloc * (y1 - y0) * self.Nx for (y0, y1) in self.geom["y"]]
        self._a2a(self.xbuf, B.view(-1), out_split, in_split)
        self.ctx.slab_pack(1, self.nloc, self.Ny, self.Nx, self.world, self.xbuf.data_ptr(), A.data_ptr())   # one scatter kernel

    # ------------------------------------------------------------------ truncated CG over the slabs
    def _poisson_cg(self, r, eps, rtol=1e-6, maxiter=1000, look=16):
        """phi (owned planes) <- scipy-cg(A, F, rtol, maxiter) as the reference calls it; returns (iterations, info)."""
        torch, dist, ctx = self.torch, self.dist, self.ctx
        L, P = self.nloc, self.P
        if self._cg is None:
            f64 = dict(dtype=torch.float64, device=self.dev)
            from .lib import lib as _lib
            self._cg = dict(r=torch.zeros((L + 2, P), **f64), p=[torch.zeros((L + 2, P), **f64) for _ in range(2)],
                            q=torch.empty((L, P), **f64), state=torch.zeros(int(_lib().foto_slab_cg_state_words()), **f64))
        c = self._cg
        rr, q, state = c["r"], c["q"], c["state"]
        x = self.phi[1].data_ptr()
        state.zero_()

        def step(op, it=0, pold=c["p"][0], pnew=c["p"][1]):
            ctx.slab_cg(op, self.Nt, self.n0, L, self.Ny, self.Nx, r, eps, rtol, it, maxiter, self.F.data_ptr(), x, rr[1].data_ptr(),
                        pold[1].data_ptr(), pnew[1].data_ptr(), q.data_ptr(), state.data_ptr())

        def allreduce():
            if self.world > 1:
                if self.staged:
                    v = state[:1].cpu(); dist.all_reduce(v); state[:1].copy_(v)
                else:
                    dist.all_reduce(state[:1])

        step(0); allreduce()
        it = 0
        while it < maxiter:
            for _ in range(min(look, maxiter - it)):
                pold, pnew = c["p"][it & 1], c["p"][(it + 1) & 1]
                self._halo([rr])
                step(3, it, pold, pnew); allreduce()          # stop test, p update incl. halo planes, q = A p, p.q
                step(5, it, pold, pnew); allreduce()          # x, r update, r.r
                it += 1
            if float(state[4]) != 0.0:
                break
        if float(state[4]) == 0.0:
            step(7, maxiter)                                  # scipy returns (x, maxiter) without another test
        st = state.tolist()
        return int(st[5]), int(st[6])

    # ------------------------------------------------------------------ solve
    def solve(self, rho0, rhoT, r=1.0, convergence_tol=0.3, reg_epsilon=1e-3, max_it=100, poisson="dct_exact"):
        """rho0, rhoT: float64 CUDA tensors of P values, identical on every rank.
        Returns (u, v, m, info) as CUDA tensors on rank 0 (None elsewhere; info everywhere)."""
        torch, dist, ctx = self.torch, self.dist, self.ctx
        L, P, Nt, Nx, Ny = self.nloc, self.P, self.Nt, self.Nx, self.Ny
        ctx.set_stream(torch.cuda.current_stream(self.dev).cuda_stream)
        self.mu.zero_(); self.q.zero_(); self.phi.zero_()
        for j in range(L):                                    # benamou_brenier.py:193-194
            w2 = (self.n0 + j) / (Nt - 1)
            self.mu[0, 1 + j] = (1 - w2) * rho0 + w2 * rhoT
        mu0, q0 = self.mu[0, 1].data_ptr(), self.q[0, 1].data_ptr()
        if poisson not in ("dct_exact", "cg_parity"):
            raise ValueError(f"unknown Poisson back-end {poisson!r}")
        crit, trace, cg_iters = -1.0, [], []
        for it in range(int(max_it)):
            self._halo([self.mu[0], self.q[0]])
            ctx.slab_rhs(mu0, q0, self.cs, rho0.data_ptr(), rhoT.data_ptr(), r, Nt, self.n0, L, Nx, Ny, self.F.data_ptr())
            if poisson == "cg_parity":
                n_cg, cg_info = self._poisson_cg(r, reg_epsilon)
                cg_iters.append(n_cg)
                if cg_info > 0 and self.rank == 0:
                    print(f"WARNING: CG did not converge in {cg_info} iterations.")      # benamou_brenier.py:86-87
            else:
                ctx.dct_xy(self.F.data_ptr(), self.A.data_ptr(), self.tmp.data_ptr(), L, Nt, Ny, Nx, False)
                self._to_y_slabs(self.A, self.B)
                ctx.dct_t_solve(self.B.data_ptr(), self.B2.data_ptr(), Nt, Ny, Nx, self.y0, self.nyl, r, reg_epsilon)
                self._to_t_slabs(self.B2, self.A)
                ctx.dct_xy(self.A.data_ptr(), self.phi[1].data_ptr(), self.tmp.data_ptr(), L, Nt, Ny, Nx, True)
            self._halo([self.phi])
            ctx.slab_prox(self.phi[1].data_ptr(), mu0, q0, self.cs, r, Nt, self.n0, L, Nx, Ny, self.sums.data_ptr())
            sums = self.sums
            if self.world > 1:
                sums = self.sums.cpu() if self.staged else self.sums
                dist.all_reduce(sums)
            num, den = (float(x) for x in sums.tolist())
            prev, crit = crit, math.sqrt(num / (den + 1e-10))          # benamou_brenier.py:246-251
            trace.append(crit)
            if crit <= convergence_tol:
                break
            if prev >= 0 and abs(prev - crit) < 1e-5:
                break
        info = dict(n_outer=len(trace), crit=np.array(trace), cg_iters=np.array(cg_iters, dtype=np.int32))
        # trajectories need every plane of phi at arbitrary (x, y): gather the owned planes on rank 0
        own = self.phi[1:L + 1].contiguous()
        if self.world > 1:
            sizes = [(n1 - n0) * P for (n0, n1) in self.geom["t"]]
            parts = [torch.empty(s, dtype=torch.float64, device=self.dev) for s in sizes] if self.rank == 0 else None
            self._gather_uneven(own.view(-1), parts, sizes) if (self.staged or len(set(sizes)) > 1) else dist.gather(own.view(-1), parts, dst=0)
            if self.rank != 0:
                return None, None, None, info
            full = torch.cat(parts)
        else:
            full = own.view(-1)
        u, v, m = (torch.empty(P, dtype=torch.float64, device=self.dev) for _ in range(3))
        ctx.flow_dev(full.data_ptr(), Nt, Nx, Ny, u.data_ptr(), v.data_ptr(), m.data_ptr())
        torch.cuda.current_stream(self.dev).synchronize()
        return u, v, m, info

    def _gather_uneven(self, mine, parts, sizes):
        dist = self.dist
        if self.rank == 0:
            parts[0].copy_(mine)
            for g in range(1, self.world):
                buf = self.torch.empty(sizes[g], dtype=mine.dtype) if self.staged else parts[g]
                dist.recv(buf, src=g)
                if self.staged:
                    parts[g].copy_(buf)
        else:
            dist.send(mine.cpu() if self.staged else mine, dst=0)
