"""`Engine`: thin Python owner of one c4_ctx (include/c4b200.h) -- the GPU-resident game pool that replaces the
reference's process x thread x pipe runtime (oinkoink/neural/game_pool.py:15-49, inference_server.py:15-76) and runs
mcts.search (oinkoink/mcts.py:94-121) for many positions at once.  PyTorch is used for device buffers and streams only.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import EVAL_CENTRE, EVAL_EXTERNAL, EVAL_NET, RNG_INJECTED, RNG_NONE, RNG_PHILOX, MCTSConfigC, ptr

RECORD_DTYPE = np.dtype({"names": ["c0", "c1", "policy", "result_value", "search_value", "game_id", "move", "ply",
                                   "n_moves", "result"],
                         "formats": ["<u8", "<u8", ("<f4", (7,)), "<f4", "<f4", "<i4", "i1", "i1", "i1", "i1"],
                         "offsets": [0, 8, 16, 44, 48, 52, 56, 57, 58, 59], "itemsize": 64})
assert RECORD_DTYPE.itemsize == 64

RAW_NODE_DTYPE = np.dtype([("vsum", "<f8"), ("visits", "<u4"), ("meta", "<u4"), ("prior", "<f8"), ("vsel", "<f8")])
assert RAW_NODE_DTYPE.itemsize == 32
NODE_DTYPE = np.dtype([("vsum", "<f8"), ("visits", "<u4"), ("meta", "<u4"), ("prior", "<f8"), ("child_block", "<u4"),
                       ("parent", "<u4"), ("vsel", "<f8")])


def _cfg_struct(cfg):
    return MCTSConfigC(int(cfg.simulations), float(cfg.pb_c_base), float(cfg.pb_c_init),
                       float(cfg.root_dirichlet_alpha), float(cfg.root_exploration_fraction),
                       int(cfg.num_sampling_moves))


def _u64_tensor(a, device=None):
    import torch
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
    if torch.is_tensor(a):
        return a.to(device=dev, dtype=torch.int64).contiguous()
    return torch.as_tensor(np.ascontiguousarray(np.asarray(a, dtype=np.uint64)).view(np.int64)).to(dev)


def _on_device(fn):
    """run an Engine method with the engine's CUDA device current: its buffers, its stream and the C side's
    cudaSetDevice(ctx->device) then all refer to the same GPU, whatever the caller's current device is"""
    import functools

    @functools.wraps(fn)
    def wrapper(self, *a, **kw):
        with self.torch.cuda.device(self.device):
            return fn(self, *a, **kw)
    return wrapper


class Engine():
    def __init__(self, max_games, config, device=None):
        import torch
        _lib.require_gpu()
        self.torch = torch
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.dev = torch.device("cuda", self.device)
        self.max_games = int(max_games)
        self.config = config
        self.lib = _lib.load()
        h = C.c_void_p()
        cs = _cfg_struct(config)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.c4_ctx_create(self.device, self.max_games, C.byref(cs), C.byref(h)))
        self.h = h
        self.net = None
        self._rng_bufs = None
        self.n_started = 0

    def close(self):
        if getattr(self, "h", None):
            self.lib.c4_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ configuration
    @_on_device
    def set_config(self, config):
        cs = _cfg_struct(config)
        _lib.check(self.lib.c4_ctx_set_config(self.h, C.byref(cs)))
        self.config = config

    @_on_device
    def set_net(self, model):
        """model: connect4_b200.neural.model.ModelWrapper (owns a c4_net)"""
        if getattr(model, "device", None) is not None and model.device.index != self.device:
            raise _lib.C4Error("the network lives on cuda:%d, the engine on cuda:%d" % (model.device.index, self.device))
        self.net = model
        _lib.check(self.lib.c4_ctx_set_net(self.h, model.c4_net))

    @_on_device
    def set_rng(self, mode="none", seed=0, noise=None, uniform=None, record=False):
        """mode: 'none' | 'philox' | 'injected'. noise [G,42,7] / uniform [G,42] float64 (numpy or CUDA tensors)."""
        torch = self.torch
        m = {"none": RNG_NONE, "philox": RNG_PHILOX, "injected": RNG_INJECTED}[mode]
        nz = un = None
        if m == RNG_INJECTED or record:
            G = self.max_games
            nz = torch.zeros((G, 42, 7), dtype=torch.float64, device=self.dev)
            un = torch.zeros((G, 42), dtype=torch.float64, device=self.dev)
            if noise is not None:
                a = torch.as_tensor(np.asarray(noise, dtype=np.float64))
                nz[:a.shape[0], :a.shape[1]] = a.to(self.dev)
            if uniform is not None:
                a = torch.as_tensor(np.asarray(uniform, dtype=np.float64))
                un[:a.shape[0], :a.shape[1]] = a.to(self.dev)
        self._rng_bufs = (nz, un)
        _lib.check(self.lib.c4_ctx_set_rng(self.h, m, int(seed) & 0xFFFFFFFFFFFFFFFF, ptr(nz), ptr(un), int(bool(record))))

    def recorded_rng(self):
        nz, un = self._rng_bufs
        return nz.cpu().numpy(), un.cpu().numpy()

    # ------------------------------------------------------------------ stand-alone searches
    @_on_device
    def begin(self, c0, c1):
        t0, t1 = _u64_tensor(c0, self.device), _u64_tensor(c1, self.device)
        n = int(t0.numel())
        _lib.check(self.lib.c4_search_begin(self.h, ptr(t0), ptr(t1), n, _lib.stream_ptr()))
        self.n_started = n
        self._roots = (t0, t1)

    @_on_device
    def run(self, kind):
        k = {"centre": EVAL_CENTRE, "net": EVAL_NET}[kind]
        _lib.check(self.lib.c4_search_run(self.h, k, _lib.stream_ptr()))

    @_on_device
    def run_external(self, evaluate_batch):
        """evaluate_batch(c0 uint64[m], c1 uint64[m]) -> (values float64[m], priors [m,7] float64 or float32).
        GPU does select / expand / backup; the callable is the reference's evaluator protocol, batched."""
        torch = self.torch
        G = self.max_games
        l0 = torch.empty(G, dtype=torch.int64, device=self.dev)
        l1 = torch.empty(G, dtype=torch.int64, device=self.dev)
        lg = torch.empty(G, dtype=torch.int32, device=self.dev)
        m = C.c_int32(0)
        while True:
            _lib.check(self.lib.c4_search_pending(self.h, ptr(l0), ptr(l1), ptr(lg), C.byref(m), _lib.stream_ptr()))
            if m.value == 0:
                break
            a = l0[:m.value].cpu().numpy().view(np.uint64)
            b = l1[:m.value].cpu().numpy().view(np.uint64)
            values, priors = evaluate_batch(a, b)
            v = torch.as_tensor(np.ascontiguousarray(values, dtype=np.float64)).to(self.dev)
            pr = np.ascontiguousarray(priors)
            dt = 1 if pr.dtype == np.float32 else 0
            if dt == 0:
                pr = pr.astype(np.float64)
            p = torch.as_tensor(pr).to(self.dev)
            _lib.check(self.lib.c4_search_supply(self.h, ptr(v), ptr(p), dt, m.value, _lib.stream_ptr()))

    @_on_device
    def readout(self, n=None):
        torch = self.torch
        n = self.n_started if n is None else n
        dev = self.dev
        out = dict(visits=torch.zeros((n, 7), dtype=torch.int32, device=dev),
                   vsum=torch.zeros((n, 7), dtype=torch.float64, device=dev),
                   cres=torch.zeros((n, 7), dtype=torch.int8, device=dev),
                   root_visits=torch.zeros(n, dtype=torch.int32, device=dev),
                   root_vsum=torch.zeros(n, dtype=torch.float64, device=dev),
                   root_prior=torch.zeros((n, 7), dtype=torch.float64, device=dev),
                   vpolicy=torch.zeros((n, 7), dtype=torch.float64, device=dev),
                   cpolicy=torch.zeros((n, 7), dtype=torch.float64, device=dev),
                   best=torch.zeros(n, dtype=torch.int8, device=dev),
                   best_value=torch.zeros(n, dtype=torch.float64, device=dev),
                   nodes=torch.zeros(n, dtype=torch.int32, device=dev))
        order = ["visits", "vsum", "cres", "root_visits", "root_vsum", "root_prior", "vpolicy", "cpolicy", "best",
                 "best_value", "nodes"]
        _lib.check(self.lib.c4_search_readout(self.h, n, *[ptr(out[k]) for k in order], _lib.stream_ptr()))
        torch.cuda.synchronize()
        return {k: v.cpu().numpy() for k, v in out.items()}

    @_on_device
    def export_tree(self, game):
        """node pool of one game as NODE_DTYPE records (the device packs the child block index into `meta` and keeps the
        side-relative value select reads next to the prior; see c4_search_export_tree in include/c4b200.h)"""
        cap = (int(self.config.simulations) + 2) * 8
        raw = np.zeros(cap, dtype=RAW_NODE_DTYPE)
        n = C.c_int64(0)
        _lib.check(self.lib.c4_search_export_tree(self.h, int(game), raw.ctypes.data_as(C.c_void_p), cap, C.byref(n),
                                                  _lib.stream_ptr()))
        raw = raw[:n.value]
        out = np.zeros(n.value, dtype=NODE_DTYPE)
        out["vsum"], out["visits"], out["prior"] = raw["vsum"], raw["visits"], raw["prior"]
        out["meta"] = raw["meta"] & 15
        out["child_block"] = raw["meta"] >> 4
        out["vsel"] = raw["vsel"]
        hdr = np.arange(n.value) % 8 == 7                       # block headers: children count / parent slot
        packed = raw["vsel"].view(np.uint64)
        out["child_block"][hdr] = (packed[hdr] & 0xFFFFFFFF).astype(np.uint32)
        out["parent"][hdr] = (packed[hdr] >> 32).astype(np.uint32)
        return out

    # ------------------------------------------------------------------ self-play
    @_on_device
    def selfplay(self, n_games, kind, game_id_base=0, game_id_stride=1, start=None, to_host=True):
        """Play n_games complete games; returns the position records (numpy structured array, RECORD_DTYPE), or with
        to_host=False leaves them on the device (`last_records_device`) and returns their number."""
        torch = self.torch
        k = {"centre": EVAL_CENTRE, "net": EVAL_NET}[kind]
        cap = int(n_games) * 42
        rec = torch.zeros((max(cap, 1), 64), dtype=torch.uint8, device=self.dev)
        s0 = s1 = None
        if start is not None:
            s0, s1 = _u64_tensor(start[0], self.device), _u64_tensor(start[1], self.device)
        n = C.c_int64(0)
        _lib.check(self.lib.c4_selfplay_run(self.h, k, int(n_games), int(game_id_base), int(game_id_stride),
                                            ptr(s0), ptr(s1), ptr(rec), cap, C.byref(n), _lib.stream_ptr()))
        self.last_records_device = rec[:n.value]
        if not to_host:
            return n.value
        return records_to_host(rec[:n.value])

    @_on_device
    def bench(self, iterations, kind):
        k = {"centre": EVAL_CENTRE, "net": EVAL_NET}[kind]
        pos, ev, sims, games = C.c_int64(0), C.c_int64(0), C.c_int64(0), C.c_int64(0)
        ms, nms, tms = C.c_float(0), C.c_float(0), C.c_float(0)
        _lib.check(self.lib.c4_selfplay_bench(self.h, k, int(iterations), C.byref(pos), C.byref(ev), C.byref(sims),
                                              C.byref(games), C.byref(ms), C.byref(nms), C.byref(tms),
                                              _lib.stream_ptr()))
        pools = self.lib.c4_ctx_get(self.h, 0) if k == EVAL_NET else 1
        return dict(positions=pos.value, evals=ev.value, sims=sims.value, games=games.value, device_ms=ms.value,
                    net_ms=nms.value, tree_ms=tms.value, iterations=int(iterations), pools=pools,
                    net_ctas=self.lib.c4_ctx_get(self.h, 1), memo_log2=self.lib.c4_ctx_get(self.h, 4),
                    memo_hits=self.lib.c4_ctx_get(self.h, 5))

    @_on_device
    def stream(self, kind, stop_games=0, max_ms=0.0, reset=False, cold_memo=False):
        """continuous self-play on the re-seeding pool (c4_selfplay_stream): until `stop_games` more games have finished
        or `max_ms` device milliseconds have passed; cold_memo empties the evaluation memo first."""
        k = {"centre": EVAL_CENTRE, "net": EVAL_NET}[kind]
        if cold_memo:
            self.clear_memo()
        pos, ev, hits, games = C.c_int64(0), C.c_int64(0), C.c_int64(0), C.c_int64(0)
        ms, eng = C.c_float(0), C.c_int32(0)
        _lib.check(self.lib.c4_selfplay_stream(self.h, k, int(bool(reset)), int(stop_games), float(max_ms), C.byref(pos),
                                               C.byref(ev), C.byref(hits), C.byref(games), C.byref(ms), C.byref(eng),
                                               _lib.stream_ptr()))
        return dict(positions=pos.value, evals=ev.value, memo_hits=hits.value, games=games.value, device_ms=ms.value,
                    engine={1: "lockstep", 2: "fused", 3: "split"}.get(eng.value, "?"), memo_log2=self.lib.c4_ctx_get(self.h, 4),
                    launches=self.lib.c4_ctx_get(self.h, 6), tree_ms=self.lib.c4_ctx_get(self.h, 8) / 1e6,
                    net_ms=self.lib.c4_ctx_get(self.h, 9) / 1e6, passes=self.lib.c4_ctx_get(self.h, 10))

    @_on_device
    def clear_memo(self):
        """a new generation starts with an empty evaluation memo (oinkoink/neural/game_pool.py:21-27)"""
        _lib.check(self.lib.c4_ctx_clear_memo(self.h, _lib.stream_ptr()))

    @_on_device
    def reset_pool(self):
        _lib.check(self.lib.c4_selfplay_reset(self.h, _lib.stream_ptr()))


def records_to_host(rec):
    """records (uint8 tensor [n, 64]) -> numpy record array on the host.  CUDA records go through page-locked memory from
    torch's caching host allocator: the buffer of one generation is reused by the next, and the copy runs at PCIe speed
    instead of the ~2 GB/s of a pageable first-touch copy (267 MB per generation at 8 GPUs)."""
    import torch
    if rec.is_cuda:
        host = torch.empty(rec.shape, dtype=rec.dtype, pin_memory=True)
        host.copy_(rec)
        rec = host
    return rec.numpy().view(RECORD_DTYPE).reshape(-1)


def augment_pack(records_device):
    """records (CUDA uint8 [n,64]) -> (boards f32 [2n,3,6,7], values f32 [2n], priors f32 [2n,7]) with the reference's
    left-right flip augmentation (oinkoink/neural/pytorch/data.py:78-105)."""
    import torch
    n = int(records_device.shape[0])
    dev = records_device.device
    boards = torch.empty((2 * n, 3, 6, 7), dtype=torch.float32, device=dev)
    values = torch.empty(2 * n, dtype=torch.float32, device=dev)
    priors = torch.empty((2 * n, 7), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().c4_records_augment_pack(ptr(records_device), n, ptr(boards), ptr(values), ptr(priors),
                                                       _lib.stream_ptr()))
    return boards, values, priors
