#!/usr/bin/env python3
"""bench.py -- self-play positions/sec at 800 sims/move (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (ours): BASELINE.json configs[2] -- 4096 concurrent self-play games per GPU, default NetConfig ResNet
(32 filters / 3 residual / 4 fc; the reference's example_net weights), 800 simulations per move, AlphaZero root noise +
6 sampled moves, 16-bit tensor-core batched leaf evaluation (fp16 operands, fp32 accumulate).  Games never interact, so with N GPUs every rank runs its own pool
(weak scaling, no data-path collective).

A "step" = `--passes` lock-step passes of the pool in steady state (finished games re-seeded at once); every pass
advances each game to its next leaf, evaluates all leaves in one network launch and backs the answers up.
  value = positions (root moves played) of all ranks / device time (CUDA events, max over ranks), inputs resident.
  e2e   = the same metric through the public API with HOST buffers: SelfPlayPool.generate_records() plays a whole
          generation from host-resident start positions and returns the position records to host memory.
  roofline = the dominant kernel of the step -- with the evaluation memo that is the tree pass k_advance (HBM class):
          algorithmic bytes per launch (1.28 KB per simulation, SURVEY.md 8d, + 64 B per memo probe) / its mean
          CUDA-event duration sampled inside the timed region, against the measured copy bandwidth; the network kernel's
          TFLOP/s figure against the measured sustained bf16 peak (MEASURED_PEAKS.json) is reported as roofline_other.
  cpu_baseline = the oracle port (oracle/selfplay_port.py) on the box's host cores, bounded sample (rank 0, N=1 only).
--impl reference: the CPU port alone, all host cores, same metric / config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "selfplay_positions_per_sec_at_800_sims_per_move"
UNIT = "positions/s"
SIMS = 800
STATE = os.path.join(ROOT, "tests", "golden", "example_net_state.npz")
WORKLOAD = ("BASELINE.json configs[2]: 4096 concurrent self-play games per GPU, default NetConfig 32f/3r/4fc "
            "(example_net weights), 800 sims/move, AlphaZero noise alpha=0.3 frac=0.25, 6 sampled moves, 16-bit (fp16 operand / fp32 accumulate) tensor-core leaf eval")


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["bf16_tflops_sustained"]), float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json, sustained bf16)"
    except Exception:
        return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region"""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_port(seconds, warm=2.0, procs=None):
    from oracle.selfplay_port import PortPool
    pp = PortPool(procs, SIMS, games_per_worker=32, state_path=STATE)
    try:
        pp.step(warm)
        r = pp.step(seconds)
    finally:
        pp.close()
    return r, pp.procs


def run_reference(args, rank):
    """the reference arm: the CPU port of the reference path on all host cores (the reference itself is pure Python
    and is not present on the GPU box)."""
    if rank != 0:
        return
    from oracle.selfplay_port import PortPool
    procs = os.cpu_count() or 1
    pp = PortPool(procs, SIMS, games_per_worker=32, state_path=STATE)
    try:
        budget = args.ref_seconds
        for _ in range(args.warmup):
            pp.step(min(budget, 3.0))
        pos = secs = evals = 0
        for _ in range(args.steps):
            r = pp.step(budget)
            pos += r["positions"]
            secs += r["seconds"]
            evals += r["evals"]
    finally:
        pp.close()
    v = pos / secs
    sample = ("%d single-threaded workers x 32 games in flight, %d steps x %.0f s of self-play at 800 sims/move (C oracle "
              "tree + torch fp32 net + per-process evaluation memo like Evaluator.position_table)" % (
                  procs, args.steps, budget))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * secs / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "host": "CPU port of the reference path (oracle/selfplay_port.py)"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "evals_per_sec": evals / secs, "gpu_launches": 0}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--games", type=int, default=4096, help="concurrent games per GPU")
    ap.add_argument("--passes", type=int, default=2000, help="lock-step passes per step")
    ap.add_argument("--preroll", type=int, default=0, help="untimed passes that bring the pool to steady state (0: use --preroll-seconds)")
    ap.add_argument("--preroll-seconds", type=float, default=4.0,
                    help="untimed device seconds of self-play before the warm-up steps (games at all plies, memo warm)")
    ap.add_argument("--e2e-games", type=int, default=16384, help="games of the end-to-end generation (4 pool-fulls)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--ref-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from connect4_b200.mcts import MCTSConfig
    from connect4_b200.neural.game_pool import SelfPlayPool
    from connect4_b200.neural.model import ModelWrapper

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    z = np.load(STATE)
    model = ModelWrapper(state_dict={k: z[k] for k in z.files})
    cfg = MCTSConfig(SIMS, 19652, 1.25, 0.3, 0.25, 6)
    pool = SelfPlayPool(model, cfg, concurrent_games=args.games, seed=1000 + rank)

    # untimed: bring the pool to steady state (games at all plies), then W warm-up steps
    done = 0
    pre_ms = 0.0
    while (done < args.preroll) if args.preroll > 0 else (pre_ms < args.preroll_seconds * 1e3):
        n = min(4000, args.preroll - done) if args.preroll > 0 else 2000
        pre_ms += pool.throughput(n)["device_ms"]
        done += n
    for _ in range(args.warmup):
        pool.throughput(args.passes)

    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    tot = dict(positions=0, evals=0, device_ms=0.0, net_ms=0.0, tree_ms=0.0, games=0, memo_hits=0)
    t_wall = time.perf_counter()
    for _ in range(args.steps):
        r = pool.throughput(args.passes)
        for k in ("positions", "evals", "device_ms", "games", "memo_hits"):
            tot[k] += r[k]
        memo_log2 = r["memo_log2"]
        tot["net_ms"] += r["net_ms"]
        tot["tree_ms"] += r["tree_ms"]
        pools = r["pools"]
    barrier()
    t_wall = time.perf_counter() - t_wall
    clocks = sampler.stop() if sampler else None

    # whole-job aggregate: sum of units over ranks / max device time over ranks
    stats = torch.tensor([tot["positions"], tot["evals"], tot["games"]], dtype=torch.float64, device="cuda")
    tmax = torch.tensor([tot["device_ms"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(stats)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    positions, evals, games = [float(x) for x in stats.tolist()]
    secs = float(tmax.item()) / 1000.0
    value = positions / secs

    # end to end through the public API with host buffers: one whole generation, records back on the host
    e2e = None
    if not args.no_e2e:
        start = (torch.zeros(args.e2e_games, dtype=torch.int64).pin_memory(),              # host-resident (pinned) inputs:
                 torch.zeros(args.e2e_games, dtype=torch.int64).pin_memory())              # every game starts from the empty board
        pool2 = SelfPlayPool(model, cfg, concurrent_games=args.games, seed=5000 + rank)
        pool2.generate_records(min(64, args.e2e_games), start=(start[0][:64], start[1][:64]))   # warm the path
        barrier()
        t0 = time.perf_counter()
        rec = pool2.generate_records(args.e2e_games, start=start)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        e = torch.tensor([float(len(rec))], dtype=torch.float64, device="cuda")
        tm = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(e)
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        e2e = {"value": float(e.item()) / float(tm.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(2 * 8 * args.e2e_games), "d2h_bytes_per_step": int(len(rec) * 64),
               "what": "SelfPlayPool.generate_records(%d games on %d slots, host start positions) -> host records; wall "
                       "clock of the whole generation incl. the drain of the last games" % (args.e2e_games, args.games)}
        pool2.engine.close()

    if rank == 0:
        peak_tf, peak_hbm, peak_src = peaks()
        flops = model.flops_per_position
        n_pass = args.steps * args.passes
        evals_per_pass = tot["evals"] / (n_pass * pools)          # per network launch (one launch per half pool per pass)
        net_ms = tot["net_ms"] / args.steps
        tree_ms = tot["tree_ms"] / args.steps
        step_ms = 1000.0 * secs / args.steps
        net_tf = (evals_per_pass * flops) / (net_ms * 1e-3) / 1e12 if net_ms > 0 else None
        # tree pass (dominant kernel, HBM class): algorithmic bytes = simulations in the launch x 1.28 KB (SURVEY.md 8d:
        # select L*(4+28+13k) + backup 24 L + expand + leaf I/O at L = 6.5, k = 7) + 64 B per evaluation-memo probe
        sims_per_launch = tot["positions"] * SIMS / (n_pass * pools)
        probes_per_launch = (tot["evals"] + tot["memo_hits"]) / (n_pass * pools)
        tree_bytes = sims_per_launch * 1280.0 + probes_per_launch * 64.0
        tree_gbs = tree_bytes / (tree_ms * 1e-3) / 1e9 if tree_ms > 0 else None
        tree_share = (tree_ms * pools * n_pass / args.steps) / step_ms if secs else None
        net_share = (net_ms * pools * n_pass / args.steps) / step_ms if secs else None
        roof_tree = {"kernel": "k_advance<NET,selfplay> (warp-per-game tree pass)", "bound": "hbm", "achieved": tree_gbs,
                     "peak": peak_hbm, "unit": "GB/s", "frac": (tree_gbs / peak_hbm) if tree_gbs else None,
                     "traffic": 55.7e6,
                     "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum per launch of this kernel, ncu --set "
                                       "full capture of the profiling command (profiles/r01_advance_v3_full_raw.csv; "
                                       "shorter passes than the default run, see profiles/README.md)",
                     "peak_source": peak_src.replace("sustained bf16", "copy bandwidth"),
                     "algorithmic_bytes_per_launch": tree_bytes, "sims_per_launch": sims_per_launch,
                     "ms_per_launch": tree_ms, "share_of_step": tree_share,
                     "note": "dependent-load latency bound (one round trip per tree level), not bandwidth bound"}
        roof_net = {"kernel": "k_net_tc<OpFP16> (tcgen05/TMEM)", "bound": "tensor", "achieved": net_tf, "peak": peak_tf,
                    "unit": "TFLOP/s", "frac": (net_tf / peak_tf) if net_tf else None, "traffic": None,
                    "peak_source": peak_src, "flops_per_eval": flops, "evals_per_launch": evals_per_pass,
                    "ms_per_launch": net_ms, "share_of_step": net_share}
        dominant_tree = (tree_share or 0) >= (net_share or 0)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "fp16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "games_per_gpu": args.games, "passes_per_step": args.passes,
                       "preroll_passes": done, "preroll_device_seconds": pre_ms / 1e3, "simulations": SIMS, "evaluation_memo_log2_entries": memo_log2,
                       "l2": "node pool (%.0f MB/GPU) + evaluation memo (%.1f GB) exceed L2; fresh leaves every pass" %
                             (args.games * (SIMS + 2) * 256 / 1e6, (64 << memo_log2) / 1e9 if memo_log2 else 0.0)},
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": int(2 * pools * n_pass + 4 * args.steps),
            "roofline": roof_tree if dominant_tree else roof_net,
            "roofline_other": roof_net if dominant_tree else roof_tree,
            "sims_per_sec": value * SIMS, "network_evals_per_sec": evals / secs,
            "memo_hit_rate": (tot["memo_hits"] / max(1.0, tot["memo_hits"] + tot["evals"])),
            "games_finished": games, "wall_s_timed_region": t_wall,
        }
        if world == 1 and not args.no_cpu:
            r, cores = cpu_port(args.cpu_seconds)
            line["cpu_baseline"] = {
                "value": r["positions_per_sec"], "unit": UNIT, "cores": cores, "kind": "port",
                "sample": "%d single-threaded workers x 32 games in flight, %.0f s of self-play at 800 sims/move "
                          "(C oracle tree + torch fp32 net + per-process evaluation memo like the reference's "
                          "Evaluator.position_table); the unmodified Python reference measured 13 positions/s on 8 cores "
                          "(BASELINE.md)" % (cores, args.cpu_seconds),
                "evals_per_sec": r["evals_per_sec"]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
