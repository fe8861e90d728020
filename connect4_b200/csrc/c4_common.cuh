// c4_common.cuh -- shared device helpers: bitboard engine, node layout, error plumbing.
// Reference semantics cited per function (paths relative to the reference root).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "../../include/c4b200.h"

typedef unsigned long long u64;

// ---------------------------------------------------------------- error plumbing (host)
void c4_set_error(const std::string &msg);
#define C4_CUDA(expr)                                                                                  \
    do {                                                                                               \
        cudaError_t _e = (expr);                                                                       \
        if (_e != cudaSuccess) {                                                                       \
            c4_set_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" +  \
                         std::to_string(__LINE__) + ")");                                              \
            return -2;                                                                                 \
        }                                                                                              \
    } while (0)
#define C4_REQUIRE(cond, msg)                                                                          \
    do {                                                                                               \
        if (!(cond)) {                                                                                 \
            c4_set_error(std::string("invalid argument: ") + (msg));                                   \
            return -1;                                                                                 \
        }                                                                                              \
    } while (0)

// ---------------------------------------------------------------- device-side checks
// compute-sanitizer is closed on the GPU pool this was developed on, so the index arithmetic of the engines carries its
// own bounds checks: a build with -DC4_CHECKED (tools/build_variant.sh checked -DC4_CHECKED) traps with file:line when one
// fails; the product build compiles them away.  tools/sanitize_case.py is the workload they are run on.
#ifdef C4_CHECKED
#include <stdio.h>
#define C4_DEV_ASSERT(cond)                                                                                         \
    do {                                                                                                            \
        if (!(cond)) {                                                                                              \
            printf("C4_DEV_ASSERT failed: %s (%s:%d) block %d thread %d\n", #cond, __FILE__, __LINE__, (int)blockIdx.x, \
                   (int)threadIdx.x);                                                                               \
            __trap();                                                                                               \
        }                                                                                                           \
    } while (0)
#else
#define C4_DEV_ASSERT(cond) do { } while (0)
#endif

// ---------------------------------------------------------------- bitboard constants (oinkoink/board.py:9-32)
#define C4_W 7
#define C4_H 6
#define C4_H1 7
#define C4_SIZE 42
#define C4_COL1 0x7FULL
#define C4_BOTTOM 0x40810204081ULL
#define C4_TOP (C4_BOTTOM << C4_H)
#define C4_RES_NONE (-1)

// Board._check_terminal_position (oinkoink/board.py:173-184): shifts 6 (\), 7 (-), 8 (/), 1 (|)
__host__ __device__ __forceinline__ bool c4_has_win(u64 b)
{
    u64 y = b & (b >> 6);
    u64 hit = y & (y >> 12);
    y = b & (b >> 7);
    hit |= y & (y >> 14);
    y = b & (b >> 8);
    hit |= y & (y >> 16);
    y = b & (b >> 1);
    hit |= y & (y >> 2);
    return hit != 0;
}

__device__ __forceinline__ int c4_age(u64 c0, u64 c1) { return __popcll(c0 | c1); }

// Board.valid_moves (oinkoink/board.py:88-92,186-188) for a running game: column c playable iff its top cell
// (bit 7c+5) is empty.  (height[c] = 7c + stones; (1<<height)&TOP != 0 iff stones == 6.)
__device__ __forceinline__ int c4_legal_mask(u64 c0, u64 c1)
{
    u64 top = (c0 | c1) >> 5;                 // bit 7c <- top cell of column c
    int m = 0;
#pragma unroll
    for (int c = 0; c < 7; c++) m |= (int)((~top >> (7 * c)) & 1ULL) << c;
    return m;
}

// bit index of the next free cell of column c: height[c] (oinkoink/board.py:39-40)
__device__ __forceinline__ int c4_drop_bit(u64 occ, int c)
{
    return 7 * c + __popcll((occ >> (7 * c)) & 0x3FULL);
}

// Board.make_move (oinkoink/board.py:160-170). `age` is the age BEFORE the move. Returns the result code.
__device__ __forceinline__ int c4_drop(u64 &c0, u64 &c1, int age, int col)
{
    u64 bit = 1ULL << c4_drop_bit(c0 | c1, col);
    bool win;
    if (age & 1) { c1 ^= bit; win = c4_has_win(c1); }
    else         { c0 ^= bit; win = c4_has_win(c0); }
    age += 1;
    if (win) return (age & 1) ? 2 : 0;        // Result(age % 2): 1.0 = o_win (code 2), 0.0 = x_win (code 0)
    if (age == C4_SIZE) return 1;             // draw
    return C4_RES_NONE;
}

// result derivation of Board.from_pieces (oinkoink/board.py:56-61)
__device__ __forceinline__ int c4_result_of(u64 c0, u64 c1)
{
    if (c4_has_win(c0)) return 2;
    if (c4_has_win(c1)) return 0;
    if (c4_age(c0, c1) == C4_SIZE) return 1;
    return C4_RES_NONE;
}

// Board.flip_color (oinkoink/board.py:127-145): mirror columns c <-> 6-c
__host__ __device__ __forceinline__ u64 c4_fliplr(u64 p)
{
    u64 r = p & (C4_COL1 << 21);
#pragma unroll
    for (int c = 0; c < 3; c++) {
        int sh = 7 * (6 - 2 * c);
        r |= (p & (C4_COL1 << (7 * c))) << sh;
        r |= (p & (C4_COL1 << (7 * (6 - c)))) >> sh;
    }
    return r;
}

// evaluate_centre (oinkoink/evaluators.py:28-33,47-63): 0.5 + (sum_o grid - sum_x grid)/96.0 with
// grid[r][c] = [0,1,2,3,2,1,0][c] + [0,1,2,2,1,0][r].  Integer dot by popcounts over weight classes.
__device__ __forceinline__ int c4_centre_weight_sum(u64 b)
{
    // column weights: cols 1,5 -> 1 ; cols 2,4 -> 2 ; col 3 -> 3
    const u64 COL = 0x3FULL;
    const u64 c1m = (COL << 7) | (COL << 35);
    const u64 c2m = (COL << 14) | (COL << 28);
    const u64 c3m = (COL << 21);
    // row weights (height h from the bottom; symmetric): h=1,4 -> 1 ; h=2,3 -> 2
    const u64 r1m = (C4_BOTTOM << 1) | (C4_BOTTOM << 4);
    const u64 r2m = (C4_BOTTOM << 2) | (C4_BOTTOM << 3);
    return __popcll(b & c1m) + 2 * __popcll(b & c2m) + 3 * __popcll(b & c3m) + __popcll(b & r1m) +
           2 * __popcll(b & r2m);
}
__device__ __forceinline__ double c4_evaluate_centre(u64 c0, u64 c1)
{
    int d = c4_centre_weight_sum(c0) - c4_centre_weight_sum(c1);
    return __dadd_rn(0.5, __ddiv_rn((double)d, 96.0));
}

// ---------------------------------------------------------------- node pool layout
// 32-byte node; the (<=7) children of a node form one 256-byte block of 8 slots (slot = column, slot 7 = block header),
// so one descent level is two fully used 128-byte lines read with 128-bit loads by lanes 0..6.  Everything the PUCT
// score of a child needs is in its own record, already in the form select consumes it:
//   A = {value_sum, visit_count, meta}: meta bit0 exists, bit1 terminal, bits2-3 result code, bits 4.. = block index of
//       this node's children (0 = not evaluated yet) -- the next level's address comes out of the same 16 bytes;
//   B = {prior, vsel}: vsel = NodeData.value(side of the parent's mover) (oinkoink/tree.py:27-44): the terminal result,
//       else value_sum / visit_count, else 0.0, flipped for x -- maintained by backup (ONE division per simulation and
//       path node, in parallel lanes) instead of being recomputed by select at every level of every descent.
// Header (slot 7): A = {position value, 0, 0}; B = {0.0, (parent node id << 32) | number of children}.
struct __align__(16) C4NodeA {     // read by select, read-modify-written by backup
    double vsum;                   // SearchEvaluation.value_sum   (oinkoink/mcts.py:46-54)
    uint32_t visits;               // SearchEvaluation.visit_count (0 = search_value is None)
    uint32_t meta;                 // flags + child block index (see above)
};
struct __align__(16) C4NodeB {
    double prior;                  // parent's PositionEvaluation.prior[this column], already normalised
    double vsel;                   // side-relative value of this node as select sees it (header: packed parent / count)
};
struct __align__(32) C4Node {
    C4NodeA a;
    C4NodeB b;
};
static_assert(sizeof(C4Node) == 32, "node must be 32 bytes");

#define C4_META_EXISTS 1u
#define C4_META_TERMINAL 2u
#define C4_META_FLAGS 15u
#define C4_META_CB_SHIFT 4
#define C4_SLOTS 8

__device__ __forceinline__ uint32_t c4_make_meta(bool exists, int result)
{
    if (!exists) return 0u;
    if (result == C4_RES_NONE) return C4_META_EXISTS;
    return C4_META_EXISTS | C4_META_TERMINAL | ((uint32_t)result << 2);
}
__device__ __forceinline__ double c4_meta_value(uint32_t meta) { return (double)((meta >> 2) & 3u) * 0.5; }
__device__ __forceinline__ int c4_meta_result(uint32_t meta)
{
    return (meta & C4_META_TERMINAL) ? (int)((meta >> 2) & 3u) : C4_RES_NONE;
}
__device__ __forceinline__ uint32_t c4_meta_child_block(uint32_t meta) { return meta >> C4_META_CB_SHIFT; }
// utils.value_to_side (oinkoink/utils.py:33-34): v for o (even age to move), 1 - v for x
__device__ __forceinline__ double c4_side_value(double v_abs, int age_of_mover)
{
    return (age_of_mover & 1) ? __dsub_rn(1.0, v_abs) : v_abs;
}

// ---------------------------------------------------------------- warp helpers
__device__ __forceinline__ double shfl_d(double v, int src)
{
    return __shfl_sync(0xffffffffu, v, src);
}
__device__ __forceinline__ u64 shfl_u64(u64 v, int src)
{
    return __shfl_sync(0xffffffffu, v, src);
}

// ---------------------------------------------------------------- Philox4x32-10 (counter-based RNG)
struct Philox {
    uint32_t key[2];
    uint32_t ctr[4];
    __device__ Philox(u64 seed, u64 stream, u64 sub)
    {
        key[0] = (uint32_t)seed; key[1] = (uint32_t)(seed >> 32);
        ctr[0] = 0; ctr[1] = (uint32_t)sub; ctr[2] = (uint32_t)stream; ctr[3] = (uint32_t)(stream >> 32);
    }
    __device__ void next(uint32_t out[4])
    {
        uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
#pragma unroll
        for (int r = 0; r < 10; r++) {
            uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
            uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
            c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
            k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
        }
        out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
        ctr[0]++;
    }
    // uniform in (0,1), 53 bits
    __device__ double uniform_from(uint32_t a, uint32_t b)
    {
        u64 x = (((u64)a << 32) | b) >> 11;
        return ((double)x + 0.5) * (1.0 / 9007199254740992.0);
    }
};
