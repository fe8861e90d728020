/*
 * c4_oracle.c -- CPU restatement of the reference's self-play hot path (willis-richard/connect4, "oinkoink").
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker for the CUDA path: only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product package (connect4_b200/) never
 * imports, links or executes anything under oracle/.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function below bit-for-bit against outputs of
 * the unmodified reference run in the build container (tests/golden/generate_goldens.py): the reference's own
 * known-answer tests (tests/board_test.py, tests/player_test.py), 6.7k random-playout board states, and
 * 10k + 295 deterministic-evaluator searches (visit counts, fp64 value sums, policies, chosen moves, node counts).
 *
 * Each function cites the reference file:line it follows (paths relative to /root/reference).
 * Structure is deliberately the reference's (one heap node per board, lazy expansion on the second visit,
 * parent pointers) -- NOT the CUDA layout (eager child blocks, path arrays) -- so the two are independent.
 *
 * Build: gcc -O2 -fPIC -shared -ffp-contract=off -o oracle/_build/libc4oracle.so oracle/c4_oracle.c -lm
 * (-ffp-contract=off: Python evaluates  pb_c*prior + value  with two roundings, never as an FMA.)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef uint64_t u64;

/* ---- bitboard constants: oinkoink/board.py:9-32.  bit c*7+h, h=0 bottom, bit c*7+6 is the column sentinel ---- */
#define WIDTH 7
#define HEIGHT 6
#define H1 7
#define H2 8
#define SIZE 42
#define COL1 127ULL
#define ALL1 ((1ULL << 49) - 1)
#define BOTTOM 0x40810204081ULL   /* ALL1 / COL1 */
#define TOP (BOTTOM << HEIGHT)

/* result codes used across the whole repo: -1 = game not over, else reference Result value = code * 0.5
 * (oinkoink/utils.py:19-22: x_win = 0.0, draw = 0.5, o_win = 1.0) */
#define RES_NONE (-1)
#define RES_XWIN 0
#define RES_DRAW 1
#define RES_OWIN 2

/* oinkoink/board.py:173-184  Board._check_terminal_position */
int c4o_has_win(u64 b)
{
    u64 y = b & (b >> HEIGHT);
    if (y & (y >> (2 * HEIGHT))) return 1;      /* diagonal \  */
    y = b & (b >> H1);
    if (y & (y >> (2 * H1))) return 1;          /* horizontal  */
    y = b & (b >> H2);
    if (y & (y >> (2 * H2))) return 1;          /* diagonal /  */
    y = b & (b >> 1);
    return (y & (y >> 2)) != 0;                 /* vertical    */
}

/* height[c] = 7*c + stones in column c  (oinkoink/board.py:39-40,163) */
static int col_height(u64 c0, u64 c1, int c)
{
    return H1 * c + __builtin_popcountll(((c0 | c1) >> (H1 * c)) & COL1);
}

int c4o_age(u64 c0, u64 c1) { return __builtin_popcountll(c0 | c1); }

/* oinkoink/board.py:88-92,186-188  Board.valid_moves / _isplayable: empty set once a result is set */
int c4o_legal_mask(u64 c0, u64 c1, int result)
{
    if (result != RES_NONE) return 0;
    int age = c4o_age(c0, c1);
    u64 me = (age & 1) ? c1 : c0;
    int m = 0;
    for (int c = 0; c < WIDTH; c++)
        if (((me | (1ULL << col_height(c0, c1, c))) & TOP) == 0) m |= 1 << c;
    return m;
}

/* oinkoink/board.py:160-170  Board.make_move (no legality check, like the reference). returns the new result code */
int c4o_drop(u64 *c0, u64 *c1, int move)
{
    int age = c4o_age(*c0, *c1);
    u64 *me = (age & 1) ? c1 : c0;
    *me ^= 1ULL << col_height(*c0, *c1, move);
    int winner = c4o_has_win(*me);
    age += 1;
    if (winner) return (age % 2) ? RES_OWIN : RES_XWIN;   /* Result(age % 2): 1.0 -> o_win, 0.0 -> x_win */
    if (age == SIZE) return RES_DRAW;
    return RES_NONE;
}

/* oinkoink/board.py:56-61  result derivation inside Board.from_pieces: o-win, then x-win, then full board */
int c4o_result_of(u64 c0, u64 c1)
{
    if (c4o_has_win(c0)) return RES_OWIN;
    if (c4o_has_win(c1)) return RES_XWIN;
    if (c4o_age(c0, c1) == SIZE) return RES_DRAW;
    return RES_NONE;
}

/* oinkoink/board.py:127-145  Board.flip_color: mirror columns c <-> 6-c */
u64 c4o_fliplr(u64 p)
{
    u64 r = 0;
    for (int c = 0; c < WIDTH; c++)
        r |= ((p >> (H1 * c)) & COL1) << (H1 * (WIDTH - 1 - c));
    return r;
}

/* oinkoink/board.py:94-113  Board.symmetrical */
int c4o_symmetrical(u64 c0, u64 c1) { return c4o_fliplr(c0) == c0 && c4o_fliplr(c1) == c1; }

/* oinkoink/board.py:64-82,147-154  Board.to_array: [3][6][7] uint8, row 0 = TOP row; ch0 = 1 iff o to move */
void c4o_to_planes(u64 c0, u64 c1, uint8_t *out)
{
    int age = c4o_age(c0, c1);
    for (int r = 0; r < HEIGHT; r++)
        for (int c = 0; c < WIDTH; c++) {
            int bit = H1 * c + (HEIGHT - 1 - r);
            out[0 * 42 + r * 7 + c] = (age % 2 == 0);
            out[1 * 42 + r * 7 + c] = (c0 >> bit) & 1;
            out[2 * 42 + r * 7 + c] = (c1 >> bit) & 1;
        }
}

/* oinkoink/board.py:43-62  Board.from_pieces (colour part): planes [6][7] (row 0 = top) -> bitboard */
u64 c4o_from_plane(const uint8_t *plane)
{
    u64 b = 0;
    for (int r = 0; r < HEIGHT; r++)
        for (int c = 0; c < WIDTH; c++)
            if (plane[r * 7 + c]) b |= 1ULL << (H1 * c + (HEIGHT - 1 - r));
    return b;
}

/* oinkoink/board.py:225-243  make_random_ips / expand: all distinct non-terminal positions after n plies.
 * Writes up to cap (c0,c1) pairs (unsorted, de-duplicated); returns the count. */
static int ips_rec(u64 c0, u64 c1, int plies, u64 *out, int n, int cap)
{
    if (plies == 0) {
        for (int i = 0; i < n; i++)
            if (out[2 * i] == c0 && out[2 * i + 1] == c1) return n;
        if (n < cap) { out[2 * n] = c0; out[2 * n + 1] = c1; }
        return n + 1;
    }
    int mask = c4o_legal_mask(c0, c1, RES_NONE);
    for (int m = 0; m < WIDTH; m++) {
        if (!(mask >> m & 1)) continue;
        u64 a = c0, b = c1;
        int res = c4o_drop(&a, &b, m);
        if (res != RES_NONE) continue;      /* terminal boards have no valid moves and are never added */
        n = ips_rec(a, b, plies - 1, out, n, cap);
    }
    return n;
}
int c4o_make_random_ips(int plies, u64 *out, int cap) { return ips_rec(0, 0, plies, out, 0, cap); }

/* oinkoink/evaluators.py:28-33,47-63  evaluate_centre: 0.5 + (sum_o grid - sum_x grid) / 96.0 ;
 * grid[r][c] = [0,1,2,3,2,1,0][c] + [0,1,2,2,1,0][r]  (symmetric in r, so row orientation is irrelevant) */
double c4o_evaluate_centre(u64 c0, u64 c1)
{
    static const int colw[7] = {0, 1, 2, 3, 2, 1, 0};
    static const int roww[6] = {0, 1, 2, 2, 1, 0};
    int so = 0, sx = 0;
    for (int c = 0; c < WIDTH; c++)
        for (int h = 0; h < HEIGHT; h++) {
            int w = colw[c] + roww[h];
            so += w * (int)((c0 >> (H1 * c + h)) & 1);
            sx += w * (int)((c1 >> (H1 * c + h)) & 1);
        }
    return 0.5 + ((double)so - (double)sx) / 96.0;
}

/* ================================================================ search ================================ */

typedef struct {
    int simulations;          /* oinkoink/mcts.py:13-26  MCTSConfig */
    double pb_c_base;
    double pb_c_init;
    double alpha;             /* root_dirichlet_alpha  */
    double frac;              /* root_exploration_fraction */
    int num_sampling_moves;
} c4o_config;

typedef struct {
    u64 c0, c1;               /* NodeData.board (oinkoink/tree.py:18-25) */
    int result;               /* board.result */
    int valid;                /* NodeData.valid_moves as a 7-bit mask */
    int name;                 /* anytree Node.name = the move that led here (root: -1) */
    int parent;
    int first_child, n_children;  /* children are created together, in ascending column order (tree.py:125-129) */
    int has_position;         /* position_value is not None */
    double pos_value;
    double prior[7];
    int has_search;           /* search_value is not None */
    double value_sum;
    int visits;
} c4o_node;

typedef struct {
    c4o_config cfg;
    c4o_node *nodes;
    int n_nodes, cap;
    int side;                 /* Tree.side (tree.py:63) */
    int sims_done;
    int pending;              /* node waiting for an evaluator answer, -1 if none */
    int root_pending;
    int has_noise;
    double noise[7];
} c4o_tree;

#define C4O_NEED_EVAL 1
#define C4O_DONE 0

static int new_node(c4o_tree *t, u64 c0, u64 c1, int result, int name, int parent)
{
    if (t->n_nodes == t->cap) {
        t->cap *= 2;
        t->nodes = (c4o_node *)realloc(t->nodes, sizeof(c4o_node) * t->cap);
    }
    c4o_node *n = &t->nodes[t->n_nodes];
    memset(n, 0, sizeof(*n));
    n->c0 = c0; n->c1 = c1; n->result = result; n->name = name; n->parent = parent;
    n->valid = c4o_legal_mask(c0, c1, result);
    n->first_child = -1;
    return t->n_nodes++;
}

/* oinkoink/tree.py:61-64  Tree.__init__ ; result must be RES_NONE for a searchable root (mcts.py:102 would raise) */
c4o_tree *c4o_tree_new(const c4o_config *cfg, u64 c0, u64 c1)
{
    c4o_tree *t = (c4o_tree *)calloc(1, sizeof(c4o_tree));
    t->cfg = *cfg;
    t->cap = 64;
    t->nodes = (c4o_node *)malloc(sizeof(c4o_node) * t->cap);
    t->side = c4o_age(c0, c1) % 2;
    new_node(t, c0, c1, c4o_result_of(c0, c1), -1, -1);
    t->pending = -1;
    return t;
}

void c4o_tree_free(c4o_tree *t) { if (t) { free(t->nodes); free(t); } }

/* oinkoink/tree.py:27-44 + utils.py:33-34  NodeData.absolute_value / value(side). NAN encodes None. */
static double absolute_value(const c4o_node *n)
{
    if (n->result != RES_NONE) return n->result * 0.5;
    if (n->has_search) return n->value_sum / (double)n->visits;
    if (n->has_position) return n->pos_value;
    return NAN;
}
static double side_value(const c4o_node *n, int side)
{
    double a = absolute_value(n);
    if (isnan(a)) return 0.0;                  /* "position is unknown - assume lost" */
    return side == 0 ? a : (1.0 - a);
}

/* oinkoink/mcts.py:147-161  ucb_score */
static double ucb_score(const c4o_tree *t, const c4o_node *p, const c4o_node *c)
{
    double pb_c = log(((double)p->visits + t->cfg.pb_c_base + 1.0) / t->cfg.pb_c_base) + t->cfg.pb_c_init;
    int cv = c->has_search ? c->visits : 0;
    pb_c = pb_c * (sqrt((double)p->visits) / (double)(cv + 1));
    double prior_score = pb_c * p->prior[c->name];
    double value_score = side_value(c, c4o_age(p->c0, p->c1) % 2);
    return prior_score + value_score;
}

/* oinkoink/mcts.py:138-144  select_child: max over (score, child); equal scores fall through to
 * Node.__gt__ = compare names (tree.py:11-15), i.e. the highest column wins ties. */
static int select_child(const c4o_tree *t, int ni)
{
    const c4o_node *p = &t->nodes[ni];
    int best = -1;
    double best_s = 0;
    for (int k = 0; k < p->n_children; k++) {
        int ci = p->first_child + k;
        double s = ucb_score(t, p, &t->nodes[ci]);
        if (best < 0 || s > best_s || (s == best_s && t->nodes[ci].name > t->nodes[best].name)) {
            best = ci; best_s = s;
        }
    }
    return best;
}

/* oinkoink/tree.py:119-132  Tree.expand_node(node, 1) */
static void expand_node(c4o_tree *t, int ni)
{
    if (t->nodes[ni].result != RES_NONE) return;
    if (t->nodes[ni].n_children) return;
    int first = t->n_nodes, cnt = 0;
    for (int m = 0; m < WIDTH; m++) {
        if (!(t->nodes[ni].valid >> m & 1)) continue;
        u64 a = t->nodes[ni].c0, b = t->nodes[ni].c1;
        int res = c4o_drop(&a, &b, m);
        new_node(t, a, b, res, m, ni);           /* may realloc: re-index t->nodes each iteration */
        cnt++;
    }
    t->nodes[ni].first_child = first;
    t->nodes[ni].n_children = cnt;
}

/* oinkoink/mcts.py:197-202  normalise(valid_moves, prior), float64 array: zero the illegal entries, divide by the
 * (sequential, n<8) numpy sum */
static void normalise64(int valid, double *p)
{
    if (valid != 127) for (int i = 0; i < 7; i++) if (!(valid >> i & 1)) p[i] = 0.0;
    double s = 0.0;
    for (int i = 0; i < 7; i++) s += p[i];
    for (int i = 0; i < 7; i++) p[i] /= s;
}
/* same, float32 array (the NN evaluator's prior is float32: neural/pytorch/model.py:265-266) */
static void normalise32(int valid, float *p)
{
    if (valid != 127) for (int i = 0; i < 7; i++) if (!(valid >> i & 1)) p[i] = 0.0f;
    float s = 0.0f;
    for (int i = 0; i < 7; i++) s = s + p[i];
    for (int i = 0; i < 7; i++) p[i] = p[i] / s;
}

/* oinkoink/mcts.py:164-168  backpropagate */
static void backpropagate(c4o_tree *t, int ni, double value)
{
    while (t->nodes[ni].parent >= 0) {
        ni = t->nodes[ni].parent;
        t->nodes[ni].value_sum += value;
        t->nodes[ni].visits += 1;
    }
}

/* Root noise to be used by the next c4o_search_supply on the root (oinkoink/mcts.py:171-181: the 7 raw
 * np.random.gamma(alpha, 1, 7) draws, injected because the reference's global MT19937 stream cannot be shared). */
void c4o_tree_set_noise(c4o_tree *t, const double *gamma7)
{
    t->has_noise = 1;
    memcpy(t->noise, gamma7, sizeof(double) * 7);
}

/* oinkoink/mcts.py:107-120  the simulation loop, run until an evaluator answer is needed or all sims are done */
int c4o_search_advance(c4o_tree *t, u64 *leaf_c0, u64 *leaf_c1)
{
    while (t->sims_done < t->cfg.simulations) {
        int ni = 0;
        while (t->nodes[ni].n_children) ni = select_child(t, ni);
        if (t->nodes[ni].has_position) {          /* previously evaluated, so expand */
            expand_node(t, ni);
            ni = select_child(t, ni);
        }
        c4o_node *n = &t->nodes[ni];
        if (n->result != RES_NONE) {              /* mcts.py:125-128 terminal branch of evaluate_node */
            double value = n->result * 0.5;
            n->has_search = 1;
            n->value_sum += value;
            n->visits += 1;
            backpropagate(t, ni, value);
            t->sims_done++;
            continue;
        }
        t->pending = ni;
        *leaf_c0 = n->c0; *leaf_c1 = n->c1;
        return C4O_NEED_EVAL;
    }
    return C4O_DONE;
}

/* oinkoink/mcts.py:98-105  search(): build the tree; the root is always evaluated first */
int c4o_search_start(c4o_tree *t, u64 *leaf_c0, u64 *leaf_c1)
{
    t->pending = 0;
    t->root_pending = 1;
    *leaf_c0 = t->nodes[0].c0; *leaf_c1 = t->nodes[0].c1;
    return C4O_NEED_EVAL;
}

/* oinkoink/mcts.py:129-135 non-terminal branch of evaluate_node, then (root) add_exploration_noise mcts.py:171-181
 * or (other nodes) backpropagate.  Exactly one of prior64 / prior32 is non-NULL; the caller's array is not modified
 * (Evaluator returns deep copies: evaluators.py:25).
 *
 * dtype note: with a float32 prior the normalisation is float32 arithmetic (numpy, in place), the
 * result is widened to double for PUCT -- the numpy-1.x behaviour the reference was written against. Under numpy>=2
 * the reference degrades the PUCT score itself to float32; that environment quirk is deliberately not reproduced. */
int c4o_search_supply(c4o_tree *t, double value, const double *prior64, const float *prior32,
                      u64 *leaf_c0, u64 *leaf_c1)
{
    c4o_node *n = &t->nodes[t->pending];
    double p[7];
    float pf[7];
    if (prior64) {
        memcpy(p, prior64, sizeof(p));
        normalise64(n->valid, p);
    } else {
        memcpy(pf, prior32, sizeof(pf));
        normalise32(n->valid, pf);
        for (int i = 0; i < 7; i++) p[i] = (double)pf[i];
    }
    n->has_position = 1;
    n->pos_value = value;
    memcpy(n->prior, p, sizeof(p));
    n->has_search = 1;
    n->value_sum = 0.0 + value;
    n->visits = 1;
    if (t->root_pending) {
        t->root_pending = 0;
        if (t->cfg.alpha != 0.0 && t->cfg.frac != 0.0 && t->has_noise) {
            double nz[7];
            memcpy(nz, t->noise, sizeof(nz));
            normalise64(n->valid, nz);
            double frac = t->cfg.frac;
            for (int i = 0; i < 7; i++) {
                /* prior * (1 - frac) + noise * frac ; a float32 prior times the python float stays float32 */
                double a = prior64 ? n->prior[i] * (1.0 - frac)
                                   : (double)(pf[i] * (float)(1.0 - frac));
                n->prior[i] = a + nz[i] * frac;
            }
        }
    } else {
        backpropagate(t, t->pending, value);
        t->sims_done++;
    }
    t->pending = -1;
    return c4o_search_advance(t, leaf_c0, leaf_c1);
}

/* Whole search with the deterministic evaluator  Evaluator(evaluate_centre_with_prior)  (evaluators.py:36-38,63) */
void c4o_search_centre(c4o_tree *t)
{
    u64 a, b;
    int st = c4o_search_start(t, &a, &b);
    while (st == C4O_NEED_EVAL) {
        double prior[7];
        for (int i = 0; i < 7; i++) prior[i] = 1.0 / 7.0;     /* np.ones(7) / 7 */
        st = c4o_search_supply(t, c4o_evaluate_centre(a, b), prior, NULL, &a, &b);
    }
}

/* ---- root read-out ---- */

/* oinkoink/tree.py:69-73  best_move: argmax of the side-relative child value, ties -> highest column */
int c4o_best_move(const c4o_tree *t)
{
    const c4o_node *r = &t->nodes[0];
    int best = -1;
    double bv = 0;
    for (int k = 0; k < r->n_children; k++) {
        const c4o_node *c = &t->nodes[r->first_child + k];
        double v = side_value(c, t->side);
        if (best < 0 || v >= bv) { best = c->name; bv = v; }
    }
    return best;
}

/* oinkoink/tree.py:75-82  sample_value_fn(lambda x: x**2) with the uniform that np.random.choice would draw:
 * p = v^2 / sum ; cdf = cumsum(p) ; cdf /= cdf[-1] ; idx = searchsorted(cdf, u, side='right') */
int c4o_sample_move(const c4o_tree *t, double u)
{
    const c4o_node *r = &t->nodes[0];
    double v[7], s = 0.0;
    int k, n = r->n_children;
    for (k = 0; k < n; k++) { v[k] = pow(side_value(&t->nodes[r->first_child + k], t->side), 2.0); }
    for (k = 0; k < n; k++) s += v[k];
    if (!(s > 0.0)) return c4o_best_move(t);     /* reference raises (NaN probabilities); documented divergence */
    double cdf[7], acc = 0.0;
    for (k = 0; k < n; k++) { acc += v[k] / s; cdf[k] = acc; }
    for (k = 0; k < n; k++) cdf[k] /= cdf[n - 1];
    int idx = 0;
    while (idx < n && cdf[idx] <= u) idx++;
    if (idx >= n) idx = n - 1;
    return t->nodes[r->first_child + idx].name;
}

/* oinkoink/tree.py:104-109,139-147  get_values_policy */
void c4o_values_policy(const c4o_tree *t, double *policy)
{
    const c4o_node *r = &t->nodes[0];
    for (int i = 0; i < 7; i++) policy[i] = 0.0;
    for (int k = 0; k < r->n_children; k++) {
        const c4o_node *c = &t->nodes[r->first_child + k];
        policy[c->name] = side_value(c, t->side);
    }
    double s = 0.0;
    for (int i = 0; i < 7; i++) s += policy[i];
    if (s == 0.0) {
        for (int k = 0; k < r->n_children; k++) policy[t->nodes[r->first_child + k].name] = 1.0;
        for (int i = 0; i < 7; i++) policy[i] /= (double)r->n_children;
    } else {
        for (int i = 0; i < 7; i++) policy[i] /= s;
    }
}

/* oinkoink/tree.py:111-117  get_visit_count_policy */
void c4o_visit_policy(const c4o_tree *t, double *policy)
{
    const c4o_node *r = &t->nodes[0];
    for (int i = 0; i < 7; i++) policy[i] = 0.0;
    for (int k = 0; k < r->n_children; k++) {
        const c4o_node *c = &t->nodes[r->first_child + k];
        if (c->has_search) policy[c->name] = (double)c->visits;
    }
    double s = 0.0;
    for (int i = 0; i < 7; i++) s += policy[i];
    if (s == 0.0) {
        for (int k = 0; k < r->n_children; k++) policy[t->nodes[r->first_child + k].name] = 1.0;
        for (int i = 0; i < 7; i++) policy[i] /= (double)r->n_children;
    } else {
        for (int i = 0; i < 7; i++) policy[i] /= s;
    }
}

/* Root statistics in column order. cres: -2 no such child, -1 non-terminal, else result code. */
void c4o_root_children(const c4o_tree *t, int32_t *visits, double *vsum, int8_t *cres, double *absval)
{
    const c4o_node *r = &t->nodes[0];
    for (int i = 0; i < 7; i++) { visits[i] = 0; vsum[i] = 0.0; cres[i] = -2; absval[i] = NAN; }
    for (int k = 0; k < r->n_children; k++) {
        const c4o_node *c = &t->nodes[r->first_child + k];
        cres[c->name] = (int8_t)c->result;
        if (c->has_search) { visits[c->name] = c->visits; vsum[c->name] = c->value_sum; }
        absval[c->name] = absolute_value(c);
    }
}
void c4o_root_stats(const c4o_tree *t, int32_t *visits, double *vsum, double *prior7, int32_t *n_nodes, int32_t *depth)
{
    *visits = t->nodes[0].visits;
    *vsum = t->nodes[0].value_sum;
    memcpy(prior7, t->nodes[0].prior, 7 * sizeof(double));
    *n_nodes = t->n_nodes;
    int maxd = 0;
    for (int i = 1; i < t->n_nodes; i++) {
        int d = 0, j = i;
        while (t->nodes[j].parent >= 0) { j = t->nodes[j].parent; d++; }
        if (d > maxd) maxd = d;
    }
    *depth = maxd;
}

/* Full node dump for whole-tree comparisons against the CUDA node pool: for every node i (creation order)
 * parent, name, result, visits (0 if never visited), value_sum. */
int c4o_tree_size(const c4o_tree *t) { return t->n_nodes; }
void c4o_tree_dump(const c4o_tree *t, int32_t *parent, int8_t *name, int8_t *result, int32_t *visits, double *vsum,
                   u64 *c0, u64 *c1)
{
    for (int i = 0; i < t->n_nodes; i++) {
        const c4o_node *n = &t->nodes[i];
        parent[i] = n->parent; name[i] = (int8_t)n->name; result[i] = (int8_t)n->result;
        visits[i] = n->has_search ? n->visits : 0;
        vsum[i] = n->has_search ? n->value_sum : 0.0;
        c0[i] = n->c0; c1[i] = n->c1;
    }
}

/* ---- batched deterministic-evaluator sweep (BASELINE.json configs[1]) ---- */
void c4o_sweep_centre(const c4o_config *cfg, int n, const u64 *c0, const u64 *c1,
                      int32_t *visits /*[n][7]*/, double *vsum /*[n][7]*/, int8_t *cres /*[n][7]*/,
                      int8_t *best, double *best_value, double *vpolicy /*[n][7]*/, int32_t *n_nodes,
                      int32_t *root_visits, double *root_vsum)
{
    for (int i = 0; i < n; i++) {
        c4o_tree *t = c4o_tree_new(cfg, c0[i], c1[i]);
        c4o_search_centre(t);
        double absval[7], prior[7];
        int32_t depth;
        c4o_root_children(t, visits + 7 * i, vsum + 7 * i, cres + 7 * i, absval);
        best[i] = (int8_t)c4o_best_move(t);
        best_value[i] = absval[best[i]];
        c4o_values_policy(t, vpolicy + 7 * i);
        c4o_root_stats(t, root_visits + i, root_vsum + i, prior, n_nodes + i, &depth);
        c4o_tree_free(t);
    }
}

/* ---- one self-play game with the deterministic evaluator (neural/training_game.py:8-19 + mcts.py:78-88) ----
 * noise: [42][7] raw gamma draws per ply or NULL; uniform: [42] one uniform per sampled ply or NULL.
 * Outputs per ply: board before the move, move, value (child.absolute_value), values-policy. Returns #plies. */
int c4o_selfplay_centre(const c4o_config *cfg, u64 start_c0, u64 start_c1, const double *noise, const double *uniform,
                        u64 *out_c0, u64 *out_c1, int8_t *out_move, double *out_value, double *out_policy,
                        int *out_result)
{
    u64 c0 = start_c0, c1 = start_c1;
    int res = c4o_result_of(c0, c1), ply = 0;
    while (res == RES_NONE) {
        c4o_tree *t = c4o_tree_new(cfg, c0, c1);
        if (noise) c4o_tree_set_noise(t, noise + 7 * ply);
        c4o_search_centre(t);
        int age = c4o_age(c0, c1), mv;
        if (age < cfg->num_sampling_moves && uniform) mv = c4o_sample_move(t, uniform[ply]);
        else mv = c4o_best_move(t);
        int32_t v[7]; double s[7], a[7]; int8_t r[7];
        c4o_root_children(t, v, s, r, a);
        out_c0[ply] = c0; out_c1[ply] = c1; out_move[ply] = (int8_t)mv; out_value[ply] = a[mv];
        c4o_values_policy(t, out_policy + 7 * ply);
        c4o_tree_free(t);
        res = c4o_drop(&c0, &c1, mv);
        ply++;
    }
    *out_result = res;
    return ply;
}
