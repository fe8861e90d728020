// c4_fz.cuh -- what the two persistent self-play engines share: the fused engine (c4_fused.cu: tree warps and a tower on
// every SM) and the split engine (c4_split.cu: tree CTAs and tower CTAs on separate SMs).  A game is run by whichever tree
// warp claims it; how the evaluator answer comes in and how a request goes out is the engine's PORT:
//   PORT::GC_MAX                                   game slots per CTA
//   int   stopping()                               the engine is draining: park the game as it is
//   float answer(g, gl, lane)                      lane < 8: {prior[7], value} of the game's answered leaf
//   void  publish(g, gl, st, request, c0, c1)      lane 0: make the game's new status visible, then (request) queue the leaf;
//                                                  st == ST_WAITMEMO: the game waits for the memo entry of (c0, c1)
//   PORT::DEDUP, bool impatient(gl)                de-duplication of evaluations in flight; a parked game asks for itself
//   PORT::GameType, void stage(G, gl)              Game, or GameS with the game's shared-memory node area filled in
#pragma once
#include "c4_tree.cuh"
#include "c4_tc.cuh"

#define FZ_WATCHDOG_CYCLES 6000000000LL                   // ~3 s without a runnable game while games wait = protocol bug
enum { FZ_ANSWERED = 5, FZ_RUNNING = 6, FZ_MEMOREADY = 8 };      // (7 = ST_WAITMEMO)

__device__ __forceinline__ unsigned long long fz_globaltimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ int ld_vol(const int *p) { return *reinterpret_cast<const volatile int *>(p); }
__device__ __forceinline__ void st_vol(int *p, int v) { *reinterpret_cast<volatile int *>(p) = v; }

// Status of a game slot at the start of a persistent launch.  A game the previous engine parked on another game's evaluation
// (ST_WAITMEMO) has applied nothing yet: it descends to that leaf again -- or, if the parked position was its root, sets the
// root up again.
__device__ __forceinline__ int fz_entry_status(const C4Dev &d, int g)
{
    const int st = d.status[g];
    if (st != ST_WAITMEMO) return st;
    return d.path_len[g] == 0 ? (int)ST_NEWROOT : (int)ST_READY;
}

// mbarrier wait that gives up when the CTA aborts (watchdog); false = aborted
__device__ __forceinline__ bool fz_wait(uint32_t bar, uint32_t parity, const int *abort)
{
    if (mbar_try(bar, parity)) return true;
    for (uint32_t it = 1;; it++) {
        if (mbar_try(bar, parity)) return true;
        if (it > 16u) __nanosleep(it > 256u ? 200 : 20);                  // an idle tower must not clog the SM's MIO queue
        if ((it & 63u) == 0u && ld_vol(abort)) return false;
    }
}

// Run game `gl` of this CTA (global slot g) until it needs the network, finishes, or the engine is stopping.
// Same state machine as k_advance (c4_search.cu), minus the pass structure.
// PORT::DEDUP (split engine): de-duplication of evaluations in flight with the PENDING tags of c4_tree.cuh -- a game whose
// leaf is being evaluated for another game parks it (ST_WAITMEMO, leaf saved like a request) and is handed back as
// FZ_MEMOREADY once the tag in the memo entry has changed; it then probes again: a hit is consumed in place of a network
// answer (bit-identical numbers), a miss (the entry was overwritten by a colliding key) claims and asks for itself.
template <bool SELFPLAY, class PORT>
__device__ __forceinline__ void fz_run_game(const C4Dev &d, const PORT &port, int g, int gl, int st, int lane)
{
    typename PORT::GameType G;
    G.g = g; G.lane = lane;
    G.gp = d.pool + (size_t)g * d.blocks_per_game * C4_SLOTS;
    port.stage(G, gl);                                                   // (split engine: the top of the tree lives in shared memory)
    G.n_blocks = d.n_blocks[g];
    G.sims_done = d.sims_done[g];
    G.c0 = d.root_c0[g]; G.c1 = d.root_c1[g];
    G.age = c4_age(G.c0, G.c1);

    bool request = false, park = false, saved = false;
    u64 rc0 = 0, rc1 = 0;
    uint32_t rnode = 0u, rlo = 0u, rhi = 0u;
    int rlen = 0;
    C4_DEV_ASSERT(gl >= 0 && gl < PORT::GC_MAX && G.n_blocks >= 1 && G.n_blocks <= d.blocks_per_game && G.sims_done <= d.sims);
    if (st == FZ_ANSWERED || (PORT::DEDUP && st == FZ_MEMOREADY)) {
        // consume the evaluator's answer for the pending leaf (oinkoink/mcts.py:129-135), then backpropagate
        const uint32_t node = (uint32_t)d.pending_node[g];
        const int plen = d.path_len[g];
        C4_DEV_ASSERT(plen >= 0 && plen <= PATH_CAP && node < (uint32_t)G.n_blocks * C4_SLOTS);
        const u64 lc0 = d.pend_c0[g], lc1 = d.pend_c1[g];
        const bool is_root = (plen == 0);
        float ov;
        bool have = true;
        if (!PORT::DEDUP || st == FZ_ANSWERED) {
            ov = port.answer(g, gl, lane);
            if (d.memo) memo_insert(d, lc0, lc1, ov, lane);
        } else {
            u64 seen = 0;
            const int pr = memo_probe(d, lc0, lc1, ov, lane, seen);
            if (pr == MEMO_HIT) {
                if (lane == 0) d.stat_hits[g] += 1ULL;
            } else {
                // still pending (the watch timed out) or overwritten by a colliding key: wait on, or ask for itself
                have = false; saved = true; rc0 = lc0; rc1 = lc1;
                if ((pr == MEMO_PENDING && !port.impatient(gl)) || (pr == MEMO_MISS && !memo_claim(d, lc0, lc1, seen, lane))) park = true;
                else request = true;
            }
        }
        if (have) {
            const double value = (double)__shfl_sync(FULL, ov, 7);
            apply_eval<true>(d, G, node, lc0, lc1, c4_age(lc0, lc1), value, 0.0, (lane < 7) ? ov : 0.f, is_root,
                             SELFPLAY ? d.ply[g] : 0);
            if (!is_root) {
                const uint32_t plo = (lane < plen) ? d.path[(size_t)g * PATH_CAP + lane] : 0u;
                const uint32_t phi = (lane + 32 < plen) ? d.path[(size_t)g * PATH_CAP + lane + 32] : 0u;
                backup(G, plo, phi, plen - 1, value);
                G.sims_done++;
            }
            st = ST_READY;
        }
    }

    while (!request && !park) {
        if (port.stopping()) break;                                    // the engine is draining: park the game as it is
        if (st == ST_NEWROOT) {
            // Tree(board) + evaluate root (oinkoink/mcts.py:98-105): fresh pool, root = node 0 of block 0
            G.n_blocks = 1;
            G.sims_done = 0;
            if (lane == 0) { G.sta(0u, 0.0, 0u, C4_META_EXISTS); G.stb(0u, 0.0, 0.0); }
            __syncwarp();
            float ov;
            u64 seen = 0;
            int pr = MEMO_MISS;
            if (d.memo) pr = PORT::DEDUP ? memo_probe(d, G.c0, G.c1, ov, lane, seen) : (memo_lookup(d, G.c0, G.c1, ov, lane) ? MEMO_HIT : MEMO_MISS);
            if (pr == MEMO_HIT) {
                apply_eval<true>(d, G, 0u, G.c0, G.c1, G.age, (double)__shfl_sync(FULL, ov, 7), 0.0, (lane < 7) ? ov : 0.f,
                                 true, SELFPLAY ? d.ply[g] : 0);
                if (lane == 0) d.stat_hits[g] += 1ULL;
                st = ST_READY;
            } else {
                rc0 = G.c0; rc1 = G.c1; rnode = 0u; rlen = 0;
                if (PORT::DEDUP && d.memo && d.memo_dedup && (pr == MEMO_PENDING || !memo_claim(d, rc0, rc1, seen, lane))) park = true;
                else request = true;
                break;
            }
        }
        if (G.sims_done >= d.sims) {
            if (!SELFPLAY) {
                st = ST_DONE;
                if (lane == 0) atomicAdd(&d.ctr->n_done, 1ULL);
                break;
            }
            st = finalize_move(d, G);
            if (st == ST_IDLE) break;
            continue;
        }
        const Leaf L = descend(d, G);
        if (L.meta & C4_META_TERMINAL) {
            // terminal branch of evaluate_node (mcts.py:125-128) + backpropagate
            backup(G, L.path_lo, L.path_hi, L.depth + 1, c4_meta_value(L.meta));
            G.sims_done++;
            continue;
        }
        u64 seen = 0;
        int pr = MEMO_MISS;
        if (d.memo) {
            float ov;
            pr = PORT::DEDUP ? memo_probe(d, L.c0, L.c1, ov, lane, seen) : (memo_lookup(d, L.c0, L.c1, ov, lane) ? MEMO_HIT : MEMO_MISS);
            if (pr == MEMO_HIT) {
                const double value = (double)__shfl_sync(FULL, ov, 7);
                apply_eval<true>(d, G, L.node, L.c0, L.c1, L.age, value, 0.0, (lane < 7) ? ov : 0.f, false, 0);
                backup(G, L.path_lo, L.path_hi, L.depth, value);
                if (lane == 0) d.stat_hits[g] += 1ULL;
                G.sims_done++;
                continue;
            }
        }
        rc0 = L.c0; rc1 = L.c1; rnode = L.node; rlen = L.depth + 1; rlo = L.path_lo; rhi = L.path_hi;
        if (PORT::DEDUP && d.memo && d.memo_dedup && (pr == MEMO_PENDING || !memo_claim(d, rc0, rc1, seen, lane))) park = true;
        else request = true;
        break;
    }
    if (request) st = ST_WAIT;
    if (park) st = ST_WAITMEMO;
    if (lane == 0) {
        d.status[g] = st;
        d.n_blocks[g] = G.n_blocks;
        d.sims_done[g] = G.sims_done;
        d.root_c0[g] = G.c0; d.root_c1[g] = G.c1;
        if ((request || park) && !saved) {
            d.pend_c0[g] = rc0; d.pend_c1[g] = rc1;
            d.pending_node[g] = (int)rnode; d.path_len[g] = rlen;
        }
        if (request) d.stat_evals[g] += 1ULL;
    }
    if ((request || park) && !saved) {
        if (lane < rlen) d.path[(size_t)g * PATH_CAP + lane] = rlo;
        if (lane + 32 < rlen) d.path[(size_t)g * PATH_CAP + lane + 32] = rhi;
    }
    __syncwarp();
    if (lane == 0) port.publish(g, gl, st, request, rc0, rc1);            // status word, then the request (if any)
    __syncwarp();
}
