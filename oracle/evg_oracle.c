/*
 * evg_oracle.c — CPU restatement of the reference's per-turn game step.  TEST INFRASTRUCTURE.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this.  It is the checker, never the product: libevgsim does not link, import or call it.
 *
 * Parity status: PINNED.  The reference ships no tests or golden vectors, so the
 * pin is the reference itself: the tests/golden npz files are trajectories of the UNMODIFIED
 * everglades_server/server.py + gym_everglades/envs/everglades_env.py run in the build
 * container under the Philox tape (oracle/ref_harness.py, tests/golden/gen_golden.py), and
 * tests/test_oracle_golden.py requires this file to reproduce them bit-for-bit (integer state,
 * observations, rewards, done flags, fp64 unit health).
 *
 * Each function cites the reference lines it restates:
 *   server.py = everglades-server/everglades_server/server.py
 *   defs.py   = everglades-server/everglades_server/definitions.py
 *   env.py    = gym-everglades/gym_everglades/envs/everglades_env.py
 * The structure is deliberately the reference's own (ordered per-node group lists, sequential
 * damage application with the nulled_ids bookkeeping); the CUDA path uses a different, parallel
 * formulation and is checked against this one.
 *
 * Third-party arithmetic on the path: numpy (unpinned by the reference; 2.3.5 here).
 *   - np.random.randint (server.py:562)  -> replaced by the tape (oracle/tape.py), both sides.
 *   - np.sum over float64 unitHealth (server.py:481) -> numpy's pairwise summation
 *     (numpy/core/src/umath/loops_utils.h.src, DOUBLE_pairwise_sum: n<8 sequential; else 8
 *     accumulators over blocks of 8, tree-combined, remainder added sequentially), restated in
 *     np_pairwise_sum() and checked against numpy in tests/test_oracle_golden.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/evgsim.h"

#define NP EVG_NUM_PLAYERS
#define NG EVG_NUM_GROUPS
#define MU EVG_MAX_GROUP_UNITS

/* ------------------------------------------------------------------ tape (oracle/tape.py) */
static void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        if (r) { k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void evo_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { philox4x32_10(ctr, key, out); }

static uint32_t tape_word(uint64_t seed, uint64_t env, uint32_t episode, uint32_t turn, uint32_t c2, uint32_t domain, int lane)
{
    uint32_t ctr[4] = {(uint32_t)env, turn, c2, domain | episode << 8};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)}, o[4];
    philox4x32_10(ctr, key, o);
    return o[lane & 3];
}

/* value the patched np.random.randint(n) returns at server.py:562 */
static uint32_t combat_draw(uint64_t seed, uint64_t env, uint32_t episode, int turn, int node, int side, int gid, int j,
                            uint32_t n)
{
    /* 8 draws of 16 bits per Philox block: word (j>>1)&3, low half for even j, high half for odd j */
    uint32_t c2 = (uint32_t)(node & 0xFF) | (uint32_t)(side & 0xFF) << 8 | (uint32_t)(gid & 0xFF) << 16 |
                  (uint32_t)((j >> 3) & 0xFF) << 24;
    uint32_t w = tape_word(seed, env, episode, (uint32_t)turn, c2, 0u, j >> 1);
    uint32_t r = (j & 1) ? w >> 16 : w & 0xFFFFu;
    return (r * n) >> 16;
}

/* ------------------------------------------------------------------ numpy float64 sum */
double evo_np_pairwise_sum(const double* a, int n)
{
    if (n < 8) {
        double res = 0.;
        for (int i = 0; i < n; ++i) res += a[i];
        return res;
    }
    /* n <= 128 on this path (group size <= 100, server.py:165), so no recursive split */
    double r[8];
    for (int k = 0; k < 8; ++k) r[k] = a[k];
    int m = n - (n % 8), i;
    for (i = 8; i < m; i += 8)
        for (int k = 0; k < 8; ++k) r[k] += a[i + k];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
}

/* ------------------------------------------------------------------ node lists
 * node.groups[pid] (defs.py:20-22) is an ordered Python list.  A group is listed at node x iff
 * location == x and it is not destroyed (appended at init server.py:198 / on arrival :691,
 * removed on departure-arrival :690 and on destruction :626-627).  The order is arrival order,
 * which EvgGroupState.arrival encodes. */
typedef struct {
    int n[EVG_MAX_NODES + 1][NP];
    int gid[EVG_MAX_NODES + 1][NP][NG];
} NodeLists;

static void build_lists(const EvgConfig* c, const EvgEnvState* s, NodeLists* L)
{
    memset(L->n, 0, sizeof(L->n));
    for (int p = 0; p < NP; ++p) {
        /* insertion sort of the player's listed groups by (arrival, gid) */
        int order[NG], m = 0;
        for (int g = 0; g < NG; ++g) {
            if (s->groups[p][g].destroyed) continue;
            int k = m++;
            while (k > 0 && s->groups[p][order[k - 1]].arrival > s->groups[p][g].arrival) {
                order[k] = order[k - 1];
                --k;
            }
            order[k] = g;
        }
        for (int k = 0; k < m; ++k) {
            int g = order[k], x = s->groups[p][g].location;
            if (x >= 1 && x <= c->n_nodes) L->gid[x][p][L->n[x][p]++] = g;
        }
    }
}

static void list_remove(NodeLists* L, int node, int p, int g)
{
    int n = L->n[node][p];
    for (int k = 0; k < n; ++k)
        if (L->gid[node][p][k] == g) {
            for (int q = k; q + 1 < n; ++q) L->gid[node][p][q] = L->gid[node][p][q + 1];
            L->n[node][p] = n - 1;
            return;
        }
}

static int alive_units(const EvgConfig* c, const EvgEnvState* s, int p, int g)
{
    int n = 0; /* np.sum(unit.unitHealth > 0), server.py:480,529 */
    for (int u = 0; u < c->group_size[p][g]; ++u) n += s->health[p][g][u] > 0;
    return n;
}

/* ------------------------------------------------------------------ capture, server.py:708-767 */
static void capture(const EvgConfig* c, EvgEnvState* s, const NodeLists* L)
{
    for (int x = 1; x <= c->n_nodes; ++x) {
        int controllers[NP], nctl = 0, points[NP];
        for (int pid = 0; pid < NP; ++pid) { /* server.py:713-726 */
            points[pid] = 0;
            int ctr = 0;
            for (int k = 0; k < L->n[x][pid]; ++k) {
                const EvgGroupState* g = &s->groups[pid][L->gid[x][pid][k]];
                if (!g->moving) {
                    ++ctr;
                    points[pid] += g->count * c->unit_control[c->group_type[pid][L->gid[x][pid][k]]];
                }
            }
            if (ctr >= 1) controllers[nctl++] = pid;
        }
        if (nctl != 1) continue; /* server.py:729 */
        int cs = s->control_state[x], cp = c->node_control_points[x];
        if (abs(cs) < cp || controllers[0] != s->controlled_by[x]) { /* server.py:731-732 */
            int pid = controllers[0] == 0 ? 0 : 1, pxer = pid == 0 ? 1 : -1;
            int neutralize = 0;
            if (s->turn == 0) { /* server.py:744-745 */
                cs = cp * pxer;
            } else { /* server.py:747-750: 0 counts as player 0's sign */
                int old_sign = cs < 0;
                cs += points[pid] * pxer;
                int new_sign = cs < 0;
                neutralize = old_sign != new_sign;
            }
            if (abs(cs) >= cp) { /* server.py:763-765 */
                cs = cp * pxer;
                s->controlled_by[x] = (int8_t)pid;
            }
            if (s->controlled_by[x] != -1 && neutralize) s->controlled_by[x] = -1; /* server.py:766-767 */
            s->control_state[x] = (int16_t)cs;
        }
    }
}

/* ------------------------------------------------------------------ game_end, server.py:281-348 */
static int game_end(const EvgConfig* c, const EvgEnvState* s, int64_t scores[NP])
{
    int base_captured[NP] = {0, 0};
    int64_t counts[NP] = {0, 0};
    scores[0] = scores[1] = 0;
    for (int x = 1; x <= c->n_nodes; ++x) { /* server.py:298-310 */
        int ts = c->node_team_start[x], cb = s->controlled_by[x], cs = s->control_state[x];
        if (ts != -1 && cb != -1 && cb != ts) {
            base_captured[ts] = 1;
            scores[cb] += c->capture_bonus;
        }
        if (cs != 0) {
            int pid = cs > 0 ? 0 : 1;
            int xer = abs(cs) == c->node_control_points[x] ? 2 : 1;
            int points = xer == 2 ? c->node_control_points[x] : abs(cs);
            scores[pid] += llabs((long long)points * xer);
        }
    }
    for (int pid = 0; pid < NP; ++pid) /* server.py:313-317 */
        for (int g = 0; g < NG; ++g)
            if (!s->groups[pid][g].destroyed) {
                counts[pid] += s->groups[pid][g].count;
                scores[pid] += (int64_t)s->groups[pid][g].count * c->unit_cost[c->group_type[pid][g]];
            }
    if (s->turn >= c->turn_limit) return EVG_STATUS_TIME_EXPIRED; /* server.py:321 */
    if (counts[0] + counts[1] == 0) return EVG_STATUS_ANNIHILATION; /* server.py:324 */
    if (base_captured[0] || base_captured[1]) return EVG_STATUS_BASE_CAPTURE; /* server.py:327 */
    return EVG_STATUS_IN_PROGRESS;
    /* server.py:337-338 draws `focus` from the global RNG every 10th turn; it is never read by
       anything on the path and the tape returns 0 for it. */
}

/* ------------------------------------------------------------------ reset
 * env.py:75-116 reset + server.py:133-209 game_init: all groups at their base in gid order,
 * health 100.0 (defs.py:62), capture() at turn 0 sets the bases to +-controlPoints. */
void evo_reset(const EvgConfig* c, EvgEnvState* s, int32_t episode)
{
    memset(s, 0, sizeof(*s));
    s->episode = episode; /* not reference state: which match of this slot, keys the tape */
    for (int x = 0; x <= EVG_MAX_NODES; ++x) s->controlled_by[x] = -1;
    for (int x = 1; x <= c->n_nodes; ++x) s->controlled_by[x] = c->node_team_start[x]; /* defs.py:16 */
    for (int p = 0; p < NP; ++p) {
        int start = -1;
        for (int x = 1; x <= c->n_nodes; ++x)
            if (c->node_team_start[x] == p) start = x; /* team_starts[p], server.py:67-68 (last wins) */
        for (int g = 0; g < NG; ++g) {
            EvgGroupState* G = &s->groups[p][g];
            G->location = (int16_t)start;
            G->travel_destination = -1;
            G->count = c->group_size[p][g];
            G->arrival = g;
            G->avg_health = 100;
            for (int u = 0; u < c->group_size[p][g]; ++u) s->health[p][g][u] = 100.;
        }
    }
    NodeLists L;
    build_lists(c, s, &L);
    capture(c, s, &L); /* server.py:206, current_turn == 0 */
}

/* Unit slots of the groups that took part in combat, summed over every evo_step of this thread: the
 * checker of the device counter EvgEpisodeStats.fought_unit_slots (16 B of health traffic per slot). */
static __thread int64_t g_fought_slots = 0;
int64_t evo_fought_slots(int clear)
{
    int64_t v = g_fought_slots;
    if (clear) g_fought_slots = 0;
    return v;
}

/* ------------------------------------------------------------------ combat, server.py:503-654 */
static void combat(const EvgConfig* c, EvgEnvState* s, NodeLists* L, uint64_t seed, uint64_t env)
{
    for (int x = 1; x <= c->n_nodes; ++x) {
        int pg[NP][NG], counts[NP][NG], npg[NP] = {0, 0}; /* player_gids, counts */
        for (int p = 0; p < NP; ++p) /* server.py:516-535 */
            for (int k = 0; k < L->n[x][p]; ++k) {
                int g = L->gid[x][p][k];
                if (!s->groups[p][g].moving) {
                    pg[p][npg[p]] = g;
                    counts[p][npg[p]] = alive_units(c, s, p, g);
                    ++npg[p];
                }
            }
        if (!(npg[0] > 0 && npg[1] > 0)) continue; /* server.py:539 */
        for (int p = 0; p < NP; ++p)
            for (int k = 0; k < npg[p]; ++k) g_fought_slots += c->group_size[p][pg[p][k]];

        /* infliction[pid][uid] += damage, server.py:549-566 */
        int infl[NP][NG * MU];
        memset(infl, 0, sizeof(infl));
        for (int pid = 0; pid < NP; ++pid) {
            int opp = 1 - pid, opp_units = 0;
            for (int k = 0; k < npg[opp]; ++k) opp_units += counts[opp][k];
            for (int i = 0; i < npg[pid]; ++i) {
                int gid = pg[pid][i], dmg = c->unit_damage[c->group_type[pid][gid]];
                for (int j = 0; j < counts[pid][i]; ++j) {
                    uint32_t uid = combat_draw(seed, env, (uint32_t)s->episode, s->turn, x, pid, gid, j, (uint32_t)opp_units);
                    infl[pid][uid] += dmg;
                }
            }
        }
        /* apply, server.py:573-643; both sides drew before anything is applied (:572) */
        int nulled[NP][NG];
        memset(nulled, 0, sizeof(nulled));
        for (int pid = 0; pid < NP; ++pid) {
            int opp = 1 - pid;
            for (int tgt0 = 0; tgt0 < NG * MU; ++tgt0) { /* sorted(infliction[pid].keys()), :578 */
                int tgt_dmg = infl[pid][tgt0];
                if (tgt_dmg == 0) continue;
                int tgt_idx = tgt0, tgt_group = 0;
                for (;;) { /* server.py:584-642 */
                    if (tgt_idx < counts[opp][tgt_group]) {
                        int tgt_gid = pg[opp][tgt_group];
                        EvgGroupState* G = &s->groups[opp][tgt_gid];
                        int type = c->group_type[opp][tgt_gid];
                        double tgt_armor = c->unit_armor[type];
                        int tgt_cntrl = s->controlled_by[x] == opp ? 1 : 0;
                        int fort_bns = c->node_has_defend[x] ? 1 : 0; /* 'DEFEND', server.py:595 */
                        double node_def = (tgt_cntrl + fort_bns) * c->node_defense[x];
                        double loss = (10. * tgt_dmg) / (tgt_armor + node_def); /* server.py:601 */
                        tgt_idx -= nulled[opp][tgt_group];                      /* server.py:605-607 */
                        /* np.argwhere(unitHealth > 0)[tgt_idx], server.py:608 */
                        int u = -1, seen = 0;
                        for (int q = 0; q < c->group_size[opp][tgt_gid]; ++q)
                            if (s->health[opp][tgt_gid][q] > 0 && seen++ == tgt_idx) { u = q; break; }
                        double h = s->health[opp][tgt_gid][u] - loss; /* server.py:609 */
                        if (h <= 0) {                                   /* server.py:615-627 */
                            h = 0;
                            G->count -= 1;
                            nulled[opp][tgt_group] += 1;
                            if (G->count == 0) {
                                G->destroyed = 1;
                                list_remove(L, x, opp, tgt_gid);
                            }
                        }
                        s->health[opp][tgt_gid][u] = h;
                        break;
                    }
                    tgt_idx -= counts[opp][tgt_group]; /* server.py:641-642 */
                    ++tgt_group;
                }
            }
        }
    }
}

/* ------------------------------------------------------------------ movement, server.py:656-706 */
static void movement(const EvgConfig* c, EvgEnvState* s, NodeLists* L)
{
    for (int p = 0; p < NP; ++p)
        for (int g = 0; g < NG; ++g) {
            EvgGroupState* G = &s->groups[p][g];
            if (G->destroyed) continue;
            if (G->ready) { /* server.py:664-667 */
                G->ready = 0;
                G->moving = 1;
            } else if (G->moving) {
                G->distance_remaining -= (int16_t)c->unit_speed[c->group_type[p][g]]; /* :671 */
                if (G->distance_remaining <= 0) {                                      /* :678-695 */
                    list_remove(L, G->location, p, g);
                    int d = G->travel_destination;
                    L->gid[d][p][L->n[d][p]++] = g;
                    G->arrival = s->turn * 16 + g;
                    G->distance_remaining = 0;
                    G->moving = 0;
                    G->location = (int16_t)d;
                    G->travel_destination = -1;
                }
            }
        }
}

/* ------------------------------------------------------------------ observations
 * board_state server.py:382-455 (fog-of-war mask :402-425 is computed there and never applied),
 * player_state server.py:457-501, concatenation env.py:158-171. */
int evo_obs_len(const EvgConfig* c) { return 1 + 4 * c->n_nodes + 5 * NG; }

void evo_observe(const EvgConfig* c, EvgEnvState* s, double* obs /* [2][obs_len] */)
{
    int len = evo_obs_len(c);
    for (int p = 0; p < NP; ++p) {
        double* o = obs + p * len;
        int opp = 1 - p, idx = 0;
        o[idx++] = s->turn;
        for (int k = 1; k <= c->n_nodes; ++k) {
            int x = p == 1 ? c->p1_node_map[k] : k; /* server.py:437-439 */
            int opp_units = 0;                      /* all listed opposing groups, moving or not, :446-449 */
            for (int g = 0; g < NG; ++g)
                if (!s->groups[opp][g].destroyed && s->groups[opp][g].location == x)
                    opp_units += s->groups[opp][g].count;
            o[idx++] = c->node_has_defense[x];
            o[idx++] = c->node_has_observe[x];
            o[idx++] = s->control_state[x];
            o[idx++] = opp_units;
        }
        for (int g = 0; g < NG; ++g) { /* server.py:475-495 */
            EvgGroupState* G = &s->groups[p][g];
            int n = c->group_size[p][g], alive = alive_units(c, s, p, g);
            double health = 0 + evo_np_pairwise_sum(s->health[p][g], n);
            int avg = alive > 0 ? (int)((health * 1.) / alive) : 0; /* int-array assignment truncates, :491 */
            G->avg_health = avg;
            o[idx++] = p == 1 ? c->p1_node_map[G->location] : G->location;
            o[idx++] = c->group_type[p][g];
            o[idx++] = avg;
            o[idx++] = G->moving ? 1 : 0;
            o[idx++] = alive;
        }
    }
}

/* ------------------------------------------------------------------ one env.step
 * env.py:32-73 step -> server.py:211-279 game_turn.  actions: int32 [2][n_rows][2].
 * Returns done (status != 0). */
int evo_step(const EvgConfig* c, EvgEnvState* s, uint64_t seed, uint64_t env, const int32_t* actions, int n_rows,
             double* obs, double reward[NP], int64_t scores[NP], int* status_out)
{
    s->turn += 1; /* server.py:214 */
    int rows = n_rows < EVG_MAX_ACTIONS ? n_rows : EVG_MAX_ACTIONS; /* action[:7,:], server.py:227 */
    for (int player = 0; player < NP; ++player) {
        int used[NG] = {0};
        for (int r = 0; r < rows; ++r) { /* server.py:232-270 */
            int gid = actions[(player * n_rows + r) * 2 + 0], nid = actions[(player * n_rows + r) * 2 + 1];
            /* Divergence (documented, include/evgsim.h): rows the reference answers with IndexError
               (gid outside [0,12), player-1 nid outside the map table) or Python negative-index
               wrap-around are no-ops here. */
            if (gid < 0 || gid >= NG) continue;
            if (player == 1) {
                if (nid < 0 || nid > c->n_nodes) continue;
                nid = c->p1_node_map[nid]; /* server.py:233-234 */
            }
            EvgGroupState* G = &s->groups[player][gid];
            int test1 = !used[gid];
            int test2 = G->moving == 0;
            int test3 = 0, distance = 0;
            if (nid >= 1 && nid <= c->n_nodes && c->edge_distance[G->location][nid]) {
                test3 = 1;
                distance = c->edge_distance[G->location][nid];
            }
            if (test1 && test2 && test3) {
                used[gid] = 1;
                G->ready = 1;
                G->moving = 0;
                G->travel_destination = (int16_t)nid;
                G->distance_remaining = (int16_t)distance;
            }
        }
    }
    NodeLists L;
    build_lists(c, s, &L);
    combat(c, s, &L, seed, env);
    movement(c, s, &L);
    capture(c, s, &L);
    /* build_knowledge_output (server.py:769-907) only formats strings that are discarded */
    int status = game_end(c, s, scores);
    if (obs) evo_observe(c, s, obs);
    /* env.py:37-60 */
    if (status != 0) {
        reward[0] = reward[1] = 0;
        if (scores[0] != scores[1]) {
            reward[0] = scores[0] > scores[1] ? 1 : 0;
            reward[1] = scores[1] > scores[0] ? 1 : -1;
        }
    } else {
        reward[0] = (double)scores[0] / c->max_score;
        reward[1] = (double)scores[1] / c->max_score;
    }
    if (status_out) *status_out = status;
    return status != 0;
}

/* ------------------------------------------------------------------ scripted agents
 * random_actions.get_action (agents/State_Machine/random_actions.py:38-46): 7 distinct group ids
 * out of 12 and 7 distinct node ids out of the map's, paired in draw order.  The reference draws
 * them with the legacy global np.random.choice; the tape version is a partial Fisher-Yates over
 * two Philox blocks: word w of block b picks position i + floor(word * (n - i) / 2^32). */
void evo_agent_random(const EvgConfig* c, uint64_t seed, uint64_t env, uint32_t episode, int turn, int player,
                      int32_t* rows /*[7][2]*/)
{
    /* 16 tape values of 16 bits: halves of the words of Philox blocks (player | 0<<8) and (player | 1<<8);
       group step k uses half k, node step k uses half 8 + k */
    uint32_t w[8], ctr[4] = {(uint32_t)env, (uint32_t)turn, 0, 1u /* DOMAIN_AGENT_RANDOM */ | episode << 8},
                   key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    for (int b = 0; b < 2; ++b) {
        ctr[2] = (uint32_t)player | (uint32_t)b << 8;
        philox4x32_10(ctr, key, w + 4 * b);
    }
    int gp[NG], np_[EVG_MAX_NODES];
    for (int i = 0; i < NG; ++i) gp[i] = i;
    for (int i = 0; i < c->n_nodes; ++i) np_[i] = i + 1;
    for (int k = 0; k < EVG_MAX_ACTIONS; ++k) {
        uint32_t hg = (k & 1) ? w[k >> 1] >> 16 : w[k >> 1] & 0xFFFFu;
        uint32_t hn = (k & 1) ? w[4 + (k >> 1)] >> 16 : w[4 + (k >> 1)] & 0xFFFFu;
        int j = k + (int)((hg * (uint32_t)(NG - k)) >> 16), t = gp[k];
        gp[k] = gp[j]; gp[j] = t;
        int nn = c->n_nodes;
        if (k < nn) {
            int q = k + (int)((hn * (uint32_t)(nn - k)) >> 16), t2 = np_[k];
            np_[k] = np_[q]; np_[q] = t2;
        }
        rows[2 * k] = gp[k];
        rows[2 * k + 1] = k < nn ? np_[k] : 0;
    }
}

/* own-numbering helpers for the observation-driven agents: an agent sees node ids through its player's map */
static int own_id(const EvgConfig* c, int player, int real) { return player == 1 ? c->p1_node_map[real] : real; }

/* base_rushV1.get_action (agents/State_Machine/base_rush_v1.py:62-111).  State word: bit0 = the blown first
 * turn is over (:73-76), bits 4-7 group_num, bits 8-15 node_num (initially 1 and 2, :55-56).  The agent keeps
 * its counters across matches, exactly like the reference object does.  Quirk kept: row i is issued when GROUP i
 * (not group_num) is away from node 11 — the enemy base in the player's own numbering (:86-88). */
void evo_agent_base_rush(const EvgConfig* c, const EvgEnvState* s, uint32_t* state, int player, int32_t* rows /*[7][2]*/)
{
    int started = *state & 1u, gnum = started ? (*state >> 4) & 15 : 1, nnum = started ? (*state >> 8) & 255 : 2;
    int enemy_base = 0;
    for (int x = 1; x <= c->n_nodes; ++x)
        if (c->node_team_start[x] == 1 - player) enemy_base = own_id(c, player, x);
    for (int i = 0; i < 2 * EVG_MAX_ACTIONS; ++i) rows[i] = 0;
    if (started) { /* act_all_cycle, :79-93 */
        for (int i = 0; i < EVG_MAX_ACTIONS; ++i) {
            if (own_id(c, player, s->groups[player][i].location) != enemy_base) {
                rows[2 * i] = gnum;
                rows[2 * i + 1] = nnum;
                gnum = (gnum + 1) % NG;
                int nodetest = nnum % c->n_nodes + 1;
                if (gnum == 0) nnum = nodetest;
            }
        }
    }
    *state = 1u | (uint32_t)gnum << 4 | (uint32_t)nnum << 8;
}

/* SwarmAgent.get_action (agents/State_Machine/swarm_agent.py:79-102).  State word: the module-level
 * ATTACK_LIST [1,2,4,5,7,8,10,11] (:24), shuffled IN PLACE on every call (:86-87) and therefore carried from
 * turn to turn, as 8 nibbles (0 = not started).  The shuffle is numpy's Fisher-Yates with the tape of
 * oracle/tape.py:swarm_shuffle.  Each listed group that is not in transit is sent to the highest-numbered
 * neighbour of where it stands (own numbering), up to 7 rows; unused rows stay [0, 1] (:81-82). */
void evo_agent_swarm(const EvgConfig* c, const EvgEnvState* s, uint32_t* state, uint64_t seed, uint64_t env, int player,
                     int32_t* rows /*[7][2]*/)
{
    uint32_t lst = *state ? *state : 0xBA875421u;
    uint32_t w[4], ctr[4] = {(uint32_t)env, (uint32_t)(s->turn + 1), (uint32_t)player, 2u | (uint32_t)s->episode << 8},
                   key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    philox4x32_10(ctr, key, w);
    for (int k = 0, i = 7; i >= 1; ++k, --i) {
        uint32_t h = (k & 1) ? w[k >> 1] >> 16 : w[k >> 1] & 0xFFFFu;
        int j = (int)((h * (uint32_t)(i + 1)) >> 16);
        uint32_t a = (lst >> (4 * i)) & 15u, b = (lst >> (4 * j)) & 15u, d = a ^ b;
        lst ^= d << (4 * i) | d << (4 * j);
    }
    *state = lst;
    for (int i = 0; i < EVG_MAX_ACTIONS; ++i) { rows[2 * i] = 0; rows[2 * i + 1] = 1; }
    int n = 0;
    for (int k = 0; k < 8 && n < EVG_MAX_ACTIONS; ++k) {
        int x = (lst >> (4 * k)) & 15;
        const EvgGroupState* G = &s->groups[player][x];
        if (!G->moving) {
            int best = 0; /* max(NODE_CONNECTIONS[pos]) in the player's own numbering */
            for (int b = 1; b <= c->n_nodes; ++b)
                if (c->edge_distance[G->location][b] && own_id(c, player, b) > best) best = own_id(c, player, b);
            rows[2 * n] = x;
            rows[2 * n + 1] = best;
            ++n;
        }
    }
}

/* ------------------------------------------------------------------ batch driver (CPU baseline)
 * Runs matches [first, first+count) for n_turns turns with both players random_actions, with
 * in-place reset on done (what the GPU arm does with EVG_AUTORESET_TERMINAL); returns the number
 * of env-turns executed and folds every observation into *checksum so nothing is optimised out. */
int64_t evo_run_random(const EvgConfig* c, uint64_t seed, int64_t first, int64_t count, int n_turns, double* checksum,
                       int64_t* episodes)
{
    double obs[2 * (1 + 4 * EVG_MAX_NODES + 5 * NG)], reward[2], acc = 0;
    int64_t scores[2], done_eps = 0, n = 0;
    int len = evo_obs_len(c);
    EvgEnvState* s = (EvgEnvState*)malloc(sizeof(EvgEnvState));
    for (int64_t e = first; e < first + count; ++e) {
        evo_reset(c, s, 0);
        for (int t = 0; t < n_turns; ++t) {
            int32_t act[2][EVG_MAX_ACTIONS][2];
            int status;
            evo_agent_random(c, seed, (uint64_t)e, (uint32_t)s->episode, s->turn + 1, 0, &act[0][0][0]);
            evo_agent_random(c, seed, (uint64_t)e, (uint32_t)s->episode, s->turn + 1, 1, &act[1][0][0]);
            int done = evo_step(c, s, seed, (uint64_t)e, &act[0][0][0], EVG_MAX_ACTIONS, obs, reward, scores, &status);
            for (int i = 0; i < 2 * len; ++i) acc += obs[i];
            acc += reward[0] + reward[1];
            ++n;
            if (done) {
                ++done_eps;
                evo_reset(c, s, s->episode + 1);
            }
        }
    }
    free(s);
    if (checksum) *checksum = acc;
    if (episodes) *episodes = done_eps;
    return n;
}

/* ------------------------------------------------------------------ lock-step batch (parity tests)
 * Steps matches first..first+n-1 (global ids) one turn each; int8 actions [n][2][7][2]; on done and
 * auto_reset != 0 the match is reset in place (episode + 1) exactly as the CUDA path does:
 * mode 1 keeps the terminal observation, mode 2 replaces it by the first observation of the new match. */
void evo_step_batch(const EvgConfig* c, EvgEnvState* states, int64_t n, uint64_t seed, int64_t first, const int8_t* actions,
                    double* obs, double* reward, uint8_t* done, int32_t* scores_out, uint8_t* status_out)
{
    int len = evo_obs_len(c);
    for (int64_t i = 0; i < n; ++i) {
        int32_t act[2 * EVG_MAX_ACTIONS * 2];
        for (int k = 0; k < 2 * EVG_MAX_ACTIONS * 2; ++k) act[k] = actions[i * 2 * EVG_MAX_ACTIONS * 2 + k];
        int64_t sc[2];
        int status;
        int d = evo_step(c, &states[i], seed, (uint64_t)(first + i), act, EVG_MAX_ACTIONS, obs + i * 2 * len, reward + 2 * i, sc,
                         &status);
        done[i] = (uint8_t)d;
        if (scores_out) { scores_out[2 * i] = (int32_t)sc[0]; scores_out[2 * i + 1] = (int32_t)sc[1]; }
        if (status_out) status_out[i] = (uint8_t)status;
        if (d && c->auto_reset != EVG_AUTORESET_OFF) {
            evo_reset(c, &states[i], states[i].episode + 1);
            if (c->auto_reset == EVG_AUTORESET_NEXT) evo_observe(c, &states[i], obs + i * 2 * len);
        }
    }
}

int evo_sizeof_config(void) { return (int)sizeof(EvgConfig); }
int evo_sizeof_state(void) { return (int)sizeof(EvgEnvState); }
