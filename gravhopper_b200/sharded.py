"""Multi-GPU leapfrog: one process per GPU (torchrun), targets sharded evenly, sources all-gathered.

SURVEY 8(e): the force on target i needs every source j, so each rank owns the state of a
contiguous slice of particles, and each step ONE collective -- an in-place NCCL all-gather of the
half-drifted positions over NVLink/NVSwitch -- gives every rank the full source array.  Direct
summation then tiles all sources against the rank's own targets; the tree is built redundantly on
every rank from the gathered sources and walked for the rank's own targets.  No reduction is
needed: each rank produces final accelerations, kicks and drifts for its own particles.

The source array is double buffered (two torch tensors bound into the engine with
gh_engine_bind_sources): while step n reads buffer b, its epilogue writes the rank's slice of
x_half(n+1) into buffer b^1, which the next all-gather completes.  Per step and rank the exchange
moves N*24 B (fp64: x_half as 3 float64) or N*16 B (fp32: float4 x_half-origin, mass).

``ShardedSimulation`` holds only host logic; the per-rank compute object (``CudaShard``) is the
C-ABI engine.  Tests drive the same host logic over gloo on CPU with a stand-in shard.
"""
import ctypes as C

import numpy as np

from . import _lib


def partition(n, world):
    """Contiguous even split of n particles over `world` ranks: list of (begin, count).
    The first n % world ranks get one extra particle."""
    if world < 1 or n < world:
        raise ValueError("need at least one particle per rank (n=%d, world=%d)" % (n, world))
    base, rem = divmod(n, world)
    out, b = [], 0
    for r in range(world):
        c = base + (1 if r < rem else 0)
        out.append((b, c))
        b += c
    return out


def morton_order(pos, bits=16):
    """Permutation that sorts particles along a 3-D Morton (Z-order) curve of their bounding box
    (host side, numpy; used only to lay particles out so that every rank's slice is spatially
    coherent -- the engines build their own exact keys on the device)."""
    pos = np.asarray(pos, dtype=np.float64)
    lo = pos.min(axis=0)
    span = np.maximum(pos.max(axis=0) - lo, 1e-300)
    q = np.minimum(((pos - lo) / span * (1 << bits)).astype(np.uint64), (1 << bits) - 1)
    key = np.zeros(len(pos), dtype=np.uint64)
    for b in range(bits):
        for k in range(3):
            key |= ((q[:, k] >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b + k)
    return np.argsort(key, kind="stable")


def interleaved_layout(pos, world, block=None):
    """Global particle order for a sharded TREE run: particles are sorted along the Morton curve,
    cut into blocks of `block`, and the blocks are dealt round-robin to the ranks; rank r's blocks
    are then stored contiguously (so ownership stays a contiguous index range).  Every warp of the
    walk then handles 32 spatial neighbours (coherent traversal), and every rank gets the same mix
    of dense and sparse regions (load balance).  Returns perm with
    new_array = old_array[perm]; the counts per rank follow `partition(n, world)`."""
    n = len(pos)
    order = morton_order(pos)
    if world == 1:
        return order
    if block is None:
        # Small blocks: measured on 8 B200 (Hernquist N = 4M) the slowest rank's walk takes 1.80 ms
        # with 2048-particle blocks and 3.42 ms with 65536-particle blocks (8 per rank): a rank that
        # holds the few densest blocks is far slower than the others, and balance matters more
        # than the ~10 % locality gain big blocks give a single rank.
        # Never fewer than ~64 blocks per rank, or the deal is too coarse to mix regions.
        block = max(32, min(2048, (n // (world * 64)) // 32 * 32))
    parts = partition(n, world)
    nblocks = (n + block - 1) // block
    owner_blocks = [[] for _ in range(world)]
    for b in range(nblocks):
        owner_blocks[b % world].append(order[b * block:(b + 1) * block])
    dealt = [np.concatenate(bl) if bl else np.zeros(0, dtype=order.dtype) for bl in owner_blocks]
    # round-robin dealing gives counts that can differ from partition() by up to one block:
    # rebalance by moving the overflow of each rank to the next one
    flat = np.concatenate(dealt)
    out, ofs = [], 0
    for (_, c) in parts:
        out.append(flat[ofs:ofs + c])
        ofs += c
    return np.concatenate(out)


class CudaShard(object):
    """One rank's engine.  native=True (default): the engine is a one-rank member of a gh_group whose
    NCCL communicator lives inside libgravhopper_b200 (ncclCommInitRank; the 128-byte id travels
    through torch.distributed) -- a step, collectives included, is ONE C call and the fp32 tree is
    built distributed.  native=False: the engine reads torch-owned source buffers and the caller
    all-gathers them with torch.distributed (redundant tree build; also what the gloo host-logic
    tests exercise with a stand-in shard)."""

    def __init__(self, n_total, begin, count, precision, device, rank=0, world=1, group=None, native=True):
        import torch
        self.torch = torch
        self.n, self.begin, self.count = n_total, begin, count
        self.precision = precision
        self.device = torch.device("cuda", device)
        self.lib = _lib.lib()
        _lib.require_gpu()
        self.native = bool(native)
        self.g = None
        prec = _lib.GH_PREC_F64 if precision == "fp64" else _lib.GH_PREC_F32
        h = C.c_void_p()
        if self.native:
            ident = (C.c_char * 128)()
            if world > 1:
                import torch.distributed as dist
                box = [None]
                if rank == 0:
                    _lib.check(self.lib.gh_group_unique_id(ident), "gh_group_unique_id")
                    box[0] = bytes(ident.raw)
                src = 0 if group is None else dist.get_global_rank(group, 0)
                dist.broadcast_object_list(box, src=src, group=group)
                ident = (C.c_char * 128).from_buffer_copy(box[0])
            g = C.c_void_p()
            _lib.check(self.lib.gh_group_create_rank(C.byref(g), ident, rank, world, device, n_total, prec),
                       "gh_group_create_rank")
            self.g = g
            b, c = C.c_int64(), C.c_int64()
            _lib.check(self.lib.gh_group_engine(g, 0, C.byref(h), C.byref(b), C.byref(c)), "gh_group_engine")
            if (b.value, c.value) != (begin, count):
                raise _lib.GravHopperB200Error("partition mismatch between the library and the host logic")
            self.h = h
            self.bufs = None
        else:
            _lib.check(self.lib.gh_engine_create(C.byref(h), device, n_total, begin, count, prec),
                       "gh_engine_create")
            self.h = h
            cols, dtype = (3, torch.float64) if precision == "fp64" else (4, torch.float32)
            self.bufs = [torch.zeros((n_total, cols), dtype=dtype, device=self.device) for _ in range(2)]
            torch.cuda.synchronize(self.device)
            _lib.check(self.lib.gh_engine_bind_sources(self.h, C.c_void_p(self.bufs[0].data_ptr()),
                                                       C.c_void_p(self.bufs[1].data_ptr())),
                       "gh_engine_bind_sources")
        s = C.c_void_p()
        _lib.check(self.lib.gh_engine_stream(self.h, C.byref(s)))
        self.stream = torch.cuda.ExternalStream(s.value, device=self.device)

    def close(self):
        h, self.h = getattr(self, "h", None), None
        g, self.g = getattr(self, "g", None), None
        try:
            if g:
                self.lib.gh_group_destroy(g)   # destroys its engine too
            elif h:
                self.lib.gh_engine_destroy(h)
        except Exception:  # interpreter shutdown: the library may already be gone
            pass

    __del__ = close

    def upload(self, pos_own, vel_own, mass_all, origin, origin_vel=None):
        pos_own = np.ascontiguousarray(pos_own, dtype=np.float64)
        vel_own = np.ascontiguousarray(vel_own, dtype=np.float64)
        mass_all = np.ascontiguousarray(mass_all, dtype=np.float64)
        origin = np.ascontiguousarray(origin, dtype=np.float64)
        _lib.check(self.lib.gh_engine_set_origin(self.h, origin.ctypes.data_as(C.POINTER(C.c_double))))
        if origin_vel is not None:
            ovel = np.ascontiguousarray(origin_vel, dtype=np.float64)
            _lib.check(self.lib.gh_engine_set_origin_velocity(self.h, ovel.ctypes.data_as(C.POINTER(C.c_double))))
        _lib.check(self.lib.gh_engine_upload(self.h, C.c_void_p(pos_own.ctypes.data),
                                             C.c_void_p(vel_own.ctypes.data),
                                             C.c_void_p(mass_all.ctypes.data)), "gh_engine_upload")

    def prepare(self, dt):
        _lib.check(self.lib.gh_engine_prepare(self.h, dt), "gh_engine_prepare")

    def source_index(self):
        i = C.c_int()
        _lib.check(self.lib.gh_engine_source_index(self.h, C.byref(i)))
        return i.value

    def step(self, dt, eps, theta, alg):
        _lib.check(self.lib.gh_engine_step(self.h, dt, eps, theta, alg, None, _lib.GH_MEM_HOST),
                   "gh_engine_step")

    def group_step(self, nsteps, dt, eps, theta, alg):
        """nsteps whole steps, NCCL exchanges included, in one asynchronous call (native only)."""
        _lib.check(self.lib.gh_group_step(self.g, int(nsteps), dt, eps, theta, alg), "gh_group_step")

    def set_tree_distributed(self, enable):
        if self.g:
            _lib.check(self.lib.gh_group_set_tree_distributed(self.g, 1 if enable else 0))

    def phase_ms(self):
        """Device ms of the last step's phases on this rank (see gh_group_phase_ms), or None."""
        if not self.g:
            return None
        out = (C.c_float * 11)()
        _lib.check(self.lib.gh_group_phase_ms(self.g, out), "gh_group_phase_ms")
        names = ("source_allgather", "build_keys_select_sort", "exchange_boundary_keys", "build_levels_scans_moments",
                 "exchange_tables", "stitch_emit", "entries_sortedindex_allgather", "walk", "acceleration_allgather",
                 "owners_kick_drift", "step")
        return {k: float(v) for k, v in zip(names, out)}

    def download(self):
        pos = np.empty((self.count, 3))
        vel = np.empty((self.count, 3))
        _lib.check(self.lib.gh_engine_download(self.h, C.c_void_p(pos.ctypes.data),
                                               C.c_void_p(vel.ctypes.data)), "gh_engine_download")
        return pos, vel

    def synchronize(self):
        _lib.check(self.lib.gh_engine_synchronize(self.h))

    def launches(self):
        n = C.c_int64()
        _lib.check(self.lib.gh_engine_launch_count(self.h, C.byref(n)))
        return n.value

    def last_force_ms(self):
        ms = C.c_float()
        _lib.check(self.lib.gh_engine_last_force_ms(self.h, C.byref(ms)))
        return ms.value

    def force_ms_mean(self, last_k):
        ms, cnt = C.c_float(), C.c_int()
        _lib.check(self.lib.gh_engine_force_ms_mean(self.h, int(last_k), C.byref(ms), C.byref(cnt)))
        return ms.value

    def stream_context(self):
        return self.torch.cuda.stream(self.stream)


class ShardedSimulation(object):
    """Leapfrog over `world` ranks.  Every rank constructs it with the FULL initial conditions
    (host arrays; synthetic ICs are generated identically on every rank) and keeps its slice."""

    def __init__(self, pos, vel, mass, dt, eps, algorithm="direct", theta=0.7, precision="fp64",
                 rank=0, world=1, device=0, group=None, shard_factory=None, layout=True, native=None):
        import torch.distributed as dist
        self.dist = dist
        self.rank, self.world, self.group = rank, world, group
        self.n = int(np.shape(pos)[0])
        self.parts = partition(self.n, world)
        self.begin, self.count = self.parts[rank]
        self.dt, self.eps, self.theta = float(dt), float(eps), float(theta)
        if algorithm not in ("direct", "tree"):
            raise ValueError("algorithm must be 'tree' or 'direct'.")
        self.alg = _lib.GH_ALG_DIRECT if algorithm == "direct" else _lib.GH_ALG_TREE
        self.uneven = len(set(c for _, c in self.parts)) > 1
        if native is None:
            import os
            native = os.environ.get("GH_COMM", "native") != "torch"
        factory = shard_factory or (lambda n, b, c: CudaShard(n, b, c, precision, device, rank=rank, world=world,
                                                              group=group, native=native))
        self.shard = factory(self.n, self.begin, self.count)
        self.native = bool(getattr(self.shard, "native", False))
        pos = np.asarray(pos, dtype=np.float64)
        vel = np.asarray(vel, dtype=np.float64)
        mass = np.asarray(mass, dtype=np.float64)
        origin = pos.mean(axis=0)  # identical on every rank: all ranks hold the same ICs
        origin_vel = vel.mean(axis=0)  # the fp32 origin moves with the system (gh_engine_set_origin_velocity)
        # tree: lay the particles out in Morton blocks dealt round-robin over the ranks (coherent
        # warps + load balance); identical on every rank.  self.perm maps new index -> original.
        self.perm = None
        if algorithm == "tree" and layout:
            self.perm = interleaved_layout(pos, world)
            pos, vel, mass = pos[self.perm], vel[self.perm], mass[self.perm]
        sl = slice(self.begin, self.begin + self.count)
        try:
            self.shard.upload(pos[sl], vel[sl], mass, origin, origin_vel)
        except TypeError:  # stand-in shards of the host-logic tests take no origin velocity
            self.shard.upload(pos[sl], vel[sl], mass, origin)
        self.shard.prepare(self.dt)
        self.steps_done = 0

    def _all_gather(self, buf):
        """In-place all-gather: every rank contributes rows [begin, begin+count) of `buf`."""
        if self.world == 1:
            return
        if not self.uneven:
            own = buf[self.begin:self.begin + self.count]
            self.dist.all_gather_into_tensor(buf, own, group=self.group)
        else:
            # uneven split (n % world != 0): one broadcast per owner; works on every backend
            for r, (b, c) in enumerate(self.parts):
                src = r if self.group is None else self.dist.get_global_rank(self.group, r)
                self.dist.broadcast(buf[b:b + c], src=src, group=self.group)

    def step(self):
        if self.native:
            self.shard.group_step(1, self.dt, self.eps, self.theta, self.alg)
        else:
            buf = self.shard.bufs[self.shard.source_index()]
            with self.shard.stream_context():
                self._all_gather(buf)
            self.shard.step(self.dt, self.eps, self.theta, self.alg)
        self.steps_done += 1

    def run(self, nsteps):
        if self.native:
            self.shard.group_step(int(nsteps), self.dt, self.eps, self.theta, self.alg)
            self.steps_done += int(nsteps)
            return
        for _ in range(int(nsteps)):
            self.step()

    def phase_ms(self):
        return self.shard.phase_ms() if hasattr(self.shard, "phase_ms") else None

    def describe(self):
        if self.world == 1:
            return "one rank: all targets and sources on one GPU"
        comm = "NCCL communicator inside libgravhopper_b200" if self.native else "torch.distributed"
        if self.alg == _lib.GH_ALG_DIRECT:
            return ("targets sharded over %d ranks, all-gather of x_half per step (%s), every rank tiles all sources"
                    % (self.world, comm))
        if self.native and self.shard.precision != "fp64":
            return ("state sharded over %d ranks, all-gather of x_half per step (%s), DISTRIBUTED tree build: every rank "
                    "sorts/scans/emits one Morton key range, two small exchanges stitch the ranges, the entry segments "
                    "and sorted indices are all-gathered, every rank walks its share of the GLOBAL Morton order "
                    "(2048-particle blocks dealt round-robin), the accelerations are all-gathered and the owners kick "
                    "and drift" % (self.world, comm))
        return ("targets sharded over %d ranks, all-gather of x_half per step (%s), tree built redundantly per rank"
                % (self.world, comm))

    def local_state(self):
        return self.shard.download()

    def close(self):
        """Release the rank's engine (device memory) now instead of at garbage collection."""
        if hasattr(self.shard, "close"):
            self.shard.close()

    def gather_state(self):
        """Full (pos, vel) on every rank (host arrays) -- output cadence only."""
        import torch
        pos, vel = self.shard.download()
        if self.world > 1:
            objs = [None] * self.world
            self.dist.all_gather_object(objs, (pos, vel), group=self.group)
            pos = np.concatenate([o[0] for o in objs], axis=0)
            vel = np.concatenate([o[1] for o in objs], axis=0)
        if self.perm is not None:  # back to the caller's particle order
            p2, v2 = np.empty_like(pos), np.empty_like(vel)
            p2[self.perm] = pos
            v2[self.perm] = vel
            pos, vel = p2, v2
        return pos, vel
