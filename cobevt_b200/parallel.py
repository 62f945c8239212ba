"""Agent-row sharding of the CoBEVT forward across the GPUs of one node.

Encoder + SinBEVT treat every agent (with its cameras) as an independent batch row, so rows
are split contiguously over ranks.  FuseBEVT is the only place agents of one sample meet:
each rank all-gathers the per-agent BEV tiles (rows x 32 x 32 x 128 16-bit, <= 256 KiB/row --
latency-bound, so a plain NCCL all-gather on the compute stream) and then runs
regroup/STTF -> FuseBEVT -> decoder -> head for the contiguous block of samples it owns.
When no sample straddles a rank boundary the exchange is skipped entirely.  ShardedCorpBEVT captures
encode + all-gather + fuse into a single CUDA graph.

The compute steps are injected (``encode`` / ``fuse``) so the same plumbing is exercised by
the world_size-2 gloo tests on CPU with the oracle's math standing in for the kernels.
"""
import torch
import torch.distributed as dist


def row_shard(total_rows, world, rank):
    """Contiguous equal split (first ranks take the remainder)."""
    base, rem = divmod(total_rows, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def sample_shard(n_samples, world, rank):
    return row_shard(n_samples, world, rank)


def needs_exchange(record_len, world):
    """True if some sample's agent rows live on more than one rank, or a rank owns a sample
    whose rows another rank encoded."""
    total = sum(record_len)
    starts = [0]
    for n in record_len:
        starts.append(starts[-1] + n)
    for r in range(world):
        r0, r1 = row_shard(total, world, r)
        s0, s1 = sample_shard(len(record_len), world, r)
        if (starts[s0], starts[s1]) != (r0, r1):
            return True
    return False


def sharded_forward(encode, fuse, local_inputs, record_len, rank, world, group=None):
    """encode(local_inputs) -> (rows_local, ...) BEV tiles for this rank's rows;
    fuse(tiles_of_owned_samples, sample_start, sample_end) -> output for the owned samples.
    ``record_len`` is the GLOBAL host-side list of agents per sample (no device sync)."""
    total = sum(record_len)
    r0, r1 = row_shard(total, world, rank)
    tiles = encode(local_inputs)
    assert tiles.shape[0] == r1 - r0, (tiles.shape, r0, r1)
    s0, s1 = sample_shard(len(record_len), world, rank)
    starts = [0]
    for n in record_len:
        starts.append(starts[-1] + n)
    if world > 1 and needs_exchange(record_len, world):
        rmax = -(-total // world)
        if tiles.shape[0] < rmax:                      # equalise for all_gather_into_tensor
            pad = torch.zeros(rmax - tiles.shape[0], *tiles.shape[1:], dtype=tiles.dtype, device=tiles.device)
            tiles = torch.cat([tiles, pad], 0)
        gathered = torch.empty(world * rmax, *tiles.shape[1:], dtype=tiles.dtype, device=tiles.device)
        dist.all_gather_into_tensor(gathered, tiles.contiguous(), group=group)
        if total % world:                              # drop the per-rank padding rows
            parts = []
            for r in range(world):
                a, b = row_shard(total, world, r)
                parts.append(gathered[r * rmax: r * rmax + (b - a)])
            gathered = torch.cat(parts, 0)
        owned = gathered[starts[s0]: starts[s1]]
    else:
        owned = tiles[starts[s0] - r0: starts[s1] - r0]
    return fuse(owned, s0, s1)


class _Graphed:
    """Replays a fixed-shape kernel sequence as one CUDA graph (launch-latency removal only: the
    captured work is exactly the eager kernel sequence).  Inputs are bound at capture time; later
    calls copy into those buffers when handed different tensors (same shape / dtype / device -- anything
    else raises).  The returned tensors are the graph's static output buffers: they are overwritten by
    the next call, clone them to keep a result."""

    def __init__(self, fn, capture_error_mode="global"):
        self.fn, self.graph, self.static_in, self.static_out = fn, None, None, None
        self.capture_error_mode = capture_error_mode

    def __call__(self, *tensors):
        if self.graph is None:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):          # warm-up outside capture: plans, smem attributes, NCCL channels
                for _ in range(2):
                    self.fn(*tensors)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self.static_in = tensors
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, capture_error_mode=self.capture_error_mode):
                self.static_out = self.fn(*self.static_in)
        else:
            if len(tensors) != len(self.static_in):
                raise ValueError(f"graph was captured with {len(self.static_in)} inputs, got {len(tensors)}")
            for i, (dst, src) in enumerate(zip(self.static_in, tensors)):
                if (dst.shape, dst.dtype, dst.device) != (src.shape, src.dtype, src.device):
                    raise ValueError(f"input {i}: graph captured for {tuple(dst.shape)} {dst.dtype} {dst.device}, "
                                     f"got {tuple(src.shape)} {src.dtype} {src.device}; build a new ShardedCorpBEVT "
                                     "for a different batch shape")
                if dst.data_ptr() != src.data_ptr():
                    dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_out


class ShardedCorpBEVT:
    """CorpBEVT (modules.py) run agent-row-sharded over a process group.

    graphs=False   every kernel launched eagerly.
    graphs=True    the whole step -- encode (encoder + SinBEVT), the NCCL all-gather of the per-agent BEV
                   tiles when a sample straddles ranks, and fuse (STTF + FuseBEVT + decoder + head) -- is
                   captured once into ONE CUDA graph (``single_graph``; NCCL collectives are capturable), so
                   no host work or graph break sits between the SinBEVT epilogue and the gather.  If the
                   capture of the collective is refused by the runtime the step falls back to two graphs with
                   an eager all-gather in between (``self.mode`` says which ran).
    The agent->ego transforms and record_len are tiny host-side values: they are re-staged into static device
    buffers on EVERY call (never cached by address), so in-place updates and recycled allocations are safe."""

    def __init__(self, model, rank=0, world=1, group=None, graphs=False, single_graph=True):
        self.m, self.rank, self.world, self.group = model, rank, world, group
        self.graphs, self.single_graph = graphs, single_graph
        self.mode = "eager" if not graphs else None
        self._g_step = self._g_encode = self._g_fuse = None
        self._layout = None
        self._rl = self._tm = None

    # ---- the two kernel sequences --------------------------------------------------------------------
    def _encode_k(self, inputs, intrinsic, extrinsic):
        return self.m.encode_nhwc(inputs, intrinsic, extrinsic)

    def _fuse_k(self, tiles, tm, rl):
        return self.m.fuse_nhwc(tiles, tm, rl)

    def _stage_layout(self, record_len, transformation_matrix, s0, s1, dev):
        """host -> static device buffers; a changed layout (different record_len slice) drops the fuse graphs."""
        key = tuple(record_len[s0:s1])
        if self._layout != key:
            self._layout = key
            self._rl = torch.tensor(list(key), dtype=torch.int64, device=dev)
            self._tm = torch.empty((s1 - s0, self.m.max_cav, 4, 4), dtype=torch.float32, device=dev)
            self._g_fuse = self._g_step = None
        self._tm.copy_(transformation_matrix[s0:s1].to(dtype=torch.float32), non_blocking=True)

    def __call__(self, local_batch, record_len, transformation_matrix):
        """local_batch: this rank's rows of inputs/intrinsic/extrinsic; record_len: global
        python list; transformation_matrix: global (B, max_cav, 4, 4) tensor (replicated)."""
        dev = local_batch['inputs'].device
        s0, s1 = sample_shard(len(record_len), self.world, self.rank)
        if s1 > s0:
            self._stage_layout(record_len, transformation_matrix, s0, s1, dev)
        lb = local_batch

        if self.graphs and self.single_graph and self.mode in (None, "single-graph") and s1 > s0:
            def step_k(inputs, intrinsic, extrinsic, tm, rl):
                return sharded_forward(lambda _: self._encode_k(inputs, intrinsic, extrinsic),
                                       lambda tiles, a, b: self._fuse_k(tiles.contiguous(), tm, rl),
                                       None, record_len, self.rank, self.world, self.group)
            if self._g_step is None:
                try:
                    g = _Graphed(step_k, capture_error_mode="thread_local")
                    out = g(lb['inputs'], lb['intrinsic'], lb['extrinsic'], self._tm, self._rl)
                    self._g_step, self.mode = g, "single-graph"
                    return out
                except Exception as e:           # capture of the collective refused: two graphs + eager gather
                    if self.world == 1:
                        raise
                    self.mode = "two-graphs"
                    self.capture_error = repr(e)[:300]
                    torch.cuda.synchronize()
            else:
                return self._g_step(lb['inputs'], lb['intrinsic'], lb['extrinsic'], self._tm, self._rl)
        if self.graphs and self.mode is None:
            self.mode = "two-graphs"

        def encode(_):
            if not self.graphs:
                return self._encode_k(lb['inputs'], lb['intrinsic'], lb['extrinsic'])
            if self._g_encode is None:
                self._g_encode = _Graphed(self._encode_k)
            return self._g_encode(lb['inputs'], lb['intrinsic'], lb['extrinsic'])

        def fuse(tiles, a, b):
            if b <= a:
                return None
            tiles = tiles.contiguous()
            if not self.graphs:
                return self._fuse_k(tiles, self._tm, self._rl)
            if self._g_fuse is None:
                self._g_fuse = _Graphed(self._fuse_k)
            return self._g_fuse(tiles, self._tm, self._rl)

        return sharded_forward(encode, fuse, None, record_len, self.rank, self.world, self.group)
