// evg_step_pair.cu — the turn step with TWO LANES PER MATCH (one per player), 128 matches per CTA.
//
// Why: the thread-per-match kernel (evg_step_tpm.cu) executes ~310 warp-instructions per match-turn but
// is latency-bound: every match keeps a 552-byte row in shared memory, so only 384 matches = 12 warps
// fit on an SM.  A match's turn is symmetric in the two players (each commands, moves and observes its
// own 12 groups), so here the two lanes of a pair split that work: same shared memory per match, same
// 384 resident matches per SM, but TWICE the warps (24) to hide shared-memory / dependency latency,
// and no divergence between the lanes of a pair because both run the same code on their own side.
//
//   lane 2k   = player 0 of the warp's match k        lane 2k+1 = player 1 of match k
//   own side:   action rows, member masks, per-node unit totals and histogram bases, movement,
//               per-node sums, unit points, reward, the player's 105 observation values, reset
//   split:      capture / node scoring (odd and even nodes), statistics (lane of player 0)
//   whole warp: record load/store (coalesced), combat (one lane per (match, fighting group) item, as in
//               evg_step_tpm.cu), observation read-out (64-byte windows, two per store instruction)
//
// Reference semantics are cited per phase (server.py / env.py as in evg_kernels.cu); the checker is
// oracle/evg_oracle.c.
#include "evg_step_common.cuh"

namespace evg {

namespace {

constexpr int kPairThreads = 256;                  // 8 warps x 16 matches
constexpr int kPairMatches = kPairThreads / 2;     // matches per CTA batch
constexpr int G12 = EVG_NUM_GROUPS;

// game_init state (server.py:133-209) of one side of one match (nodes split between the two lanes)
__device__ __noinline__ void reset_side(const Tables& S, uint32_t* R, double* health_side, int n_nodes, int sd)
{
    for (int g = 0; g < G12; ++g) {
        const int L = sd * G12 + g;
        R[2 * L] = S.init_w0[L];
        R[2 * L + 1] = S.init_w1[L];
    }
    for (int n = 1 + sd; n <= n_nodes; n += 2) R[kRecNode0 + n - 1] = S.init_node[n];
    double2* hp = reinterpret_cast<double2*>(health_side);
    for (int i = 0; i < S.health_slots / 4; ++i) hp[i] = make_double2(100.0, 100.0);  // definitions.py:62
}

template <int NODES, int MAXSZ, typename HistT, int PITCH>
__global__ void __launch_bounds__(kPairThreads, 3) evg_step_pair_kernel(const __grid_constant__ Tables T, const __grid_constant__ StepArgs A)
{
    extern __shared__ __align__(16) unsigned char smem[];
    {   // ---- stage the static tables once per CTA
        const uint32_t* src = reinterpret_cast<const uint32_t*>(&T);
        uint32_t* dst = reinterpret_cast<uint32_t*>(smem);
        for (int i = threadIdx.x; i < (int)(sizeof(Tables) / 4); i += blockDim.x) dst[i] = src[i];
    }
    const Tables& S = *reinterpret_cast<const Tables*>(smem);
    __syncthreads();

    const Geo<NODES> G(S);
    const int n_nodes = G.n_nodes(), nn = G.nn(), RW = G.rw(), OL = G.obs_len();
    const int P = PITCH ? PITCH : T.pair_pitch;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int pm = lane >> 1, sd = lane & 1;  // my match within the warp, my side
    uint32_t* rows = reinterpret_cast<uint32_t*>(smem + T.sm_tables_bytes + 128);
    uint32_t* wrow = rows + (size_t)warp * 16 * P;  // the warp's 16 rows
    uint32_t* R = wrow + (size_t)pm * P;            // my match's record
    uint32_t* X = R + RW;                           // my match's scratch
    uint16_t* X16 = reinterpret_cast<uint16_t*>(X);
    const int hslots_side = S.health_slots / 2;
    // scratch layout (words): combat  X[0..nn) member masks (u16 half per side), X[nn..2nn) per-node totals/bases
    //                                 (u16 half per side: total | base << 8), X[2nn..) damage histograms of side 0, 1
    //                         post    X[0..nn) side-0 sums, X[nn..2nn) side-1 sums, X[2nn..2nn+32) obs windows
    const int64_t nbatches = (A.n_envs + kPairMatches - 1) / kPairMatches;
    for (int64_t batch = blockIdx.x; batch < nbatches; batch += gridDim.x) {
    const int64_t warp_env0 = batch * kPairMatches + warp * 16;
    const int64_t env = warp_env0 + pm;
    const int64_t left = A.n_envs - warp_env0;
    const int nvalid = left >= 16 ? 16 : (left > 0 ? (int)left : 0);
    const bool valid = pm < nvalid;
    {   // pull the NEXT batch's records and action rows towards L2 while this one is processed
        const int64_t nenv0 = warp_env0 + (int64_t)gridDim.x * kPairMatches;
        if (nenv0 + 16 <= A.n_envs) {
            const char* nr = reinterpret_cast<const char*>(A.records) + nenv0 * RW * 4;
            for (int b = lane * 128; b < 16 * RW * 4; b += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(nr + b));
            const char* na = reinterpret_cast<const char*>(A.actions) + nenv0 * 28;
            if (lane < 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(na + lane * 128));
        }
    }

    // my player's 7 action rows (14 bytes), requested before the records so the latencies overlap
    uint32_t rows7[EVG_MAX_ACTIONS];
    const bool ext = A.agent[sd] == EVG_AGENT_EXTERNAL;
#pragma unroll
    for (int k = 0; k < EVG_MAX_ACTIONS; ++k)
        rows7[k] = (valid && ext) ? __ldcs(reinterpret_cast<const uint16_t*>(A.actions) + (env * 2 + sd) * EVG_MAX_ACTIONS + k) : 0u;

    // ---- cooperative, coalesced load of the warp's records into the per-match rows
    {
        const int q4 = RW / 4;  // 16-byte chunks per record
        const uint4* g4 = reinterpret_cast<const uint4*>(A.records) + warp_env0 * q4;
        const int total = nvalid * q4;
#pragma unroll 8
        for (int f = lane; f < total; f += 32) {
            const int m = NODES ? f >> 4 : f / q4, q = NODES ? f & 15 : f % q4;
            const uint4 v = __ldcs(g4 + f);
            uint2* d = reinterpret_cast<uint2*>(wrow + (size_t)m * P + 4 * q);
            d[0] = make_uint2(v.x, v.y);
            d[1] = make_uint2(v.z, v.w);
        }
    }
    __syncwarp();

    uint32_t turn = 0, episode = 0;
    uint32_t fm = 0;  // my side's fighting groups (bit g)
    if (valid) {
        turn = R[kRecTurn] + 1u;  // server.py:214
        episode = R[kRecEpisode];

        // ---- action decode + validation of MY player's rows, server.py:218-271 (first valid row per group wins)
        {
            if (A.agent[sd] == EVG_AGENT_RANDOM)  // agents/State_Machine/random_actions.py:38-46 on the tape
                agent_random_rows(S.env_base + (uint32_t)env, turn, episode, sd, n_nodes, S.seed_lo, S.seed_hi, rows7);
            if (A.actions_out && A.agent[sd] != EVG_AGENT_EXTERNAL) {
                uint16_t* ao = reinterpret_cast<uint16_t*>(A.actions_out) + (env * 2 + sd) * EVG_MAX_ACTIONS;
#pragma unroll
                for (int k = 0; k < EVG_MAX_ACTIONS; ++k) ao[k] = (uint16_t)rows7[k];
            }
            int Lr[EVG_MAX_ACTIONS];
            uint32_t dr[EVG_MAX_ACTIONS], nr[EVG_MAX_ACTIONS];
#pragma unroll
            for (int r = 0; r < EVG_MAX_ACTIONS; ++r) {  // all lookups first: accepting a row changes neither loc nor moving
                const uint32_t a = rows7[r];
                const int ag = (int)(int8_t)(a & 0xFFu);
                int an = (int)(int8_t)(a >> 8);
                const bool okg = (unsigned)ag < (unsigned)G12;
                an = (unsigned)an <= (unsigned)n_nodes ? an : 0;
                if (sd) an = S.p1_map[an];  // server.py:233-234
                const int L = sd * G12 + (okg ? ag : 0);
                const uint32_t gw0 = R[2 * L];
                const uint32_t d = S.edge[gw0 & W0_LOC_MASK][an];
                Lr[r] = L;
                nr[r] = (uint32_t)an;
                dr[r] = (okg && !(gw0 & W0_MOVING)) ? d : 0u;  // t2 (not moving) and t3 (adjacent), :243-250
            }
            uint32_t used = 0;
#pragma unroll
            for (int r = 0; r < EVG_MAX_ACTIONS; ++r) {
                if (dr[r] && !((used >> Lr[r]) & 1u)) {  // t1: no accepted command for the group yet, :241
                    used |= 1u << Lr[r];
                    const uint32_t gw0 = R[2 * Lr[r]];
                    R[2 * Lr[r]] = (gw0 & ~((0x3Fu << W0_DEST_SHIFT) | (0xFFu << W0_DIST_SHIFT))) | nr[r] << W0_DEST_SHIFT |
                                   dr[r] << W0_DIST_SHIFT | W0_READY;  // :267-270
                }
            }
        }

        // ---- combat preparation, my side (server.py:516-553): member masks of my present groups per node
        for (int x = 0; x < nn; ++x) X16[2 * x + sd] = 0;
#pragma unroll 4
        for (int g = 0; g < G12; ++g) {  // listed and not in transit, :516-535
            const uint2 w = *reinterpret_cast<const uint2*>(R + 2 * (sd * G12 + g));
            if ((w.y & 0xFFFFu) && !(w.x & W0_MOVING)) X16[2 * (w.x & W0_LOC_MASK) + sd] |= (uint16_t)(1u << g);
        }
    }
    __syncwarp();  // the partner's masks
    if (valid) {
        uint32_t base = 0;
        for (int x = 1; x <= n_nodes; ++x) {
            const uint32_t mine = X16[2 * x + sd];
            if (mine && X16[2 * x + 1 - sd]) {  // both players present: contested, :539
                fm |= mine;
                uint32_t t = 0;  // np.sum(counts[pid]), :552-553
                for (uint32_t m = mine; m; m &= m - 1) t += __popc(R[2 * (sd * G12 + __ffs(m) - 1) + 1] & 0xFFFFu);
                X16[2 * (nn + x) + sd] = (uint16_t)(t | base << 8);  // node-local uid -> histogram slot base
                base += t;
            }
        }
        uint32_t* H = X + 2 * nn + sd * S.pair_hwords;  // targets on my side
        for (int i = 0; i < (int)((base * sizeof(HistT) + 3) / 4); ++i) H[i] = 0;
        const char* he = reinterpret_cast<const char*>(A.health + env * S.health_slots);
        for (uint32_t m = fm; m; m &= m - 1) {  // start my fighting groups' health rows towards L2
            const int L = sd * G12 + __ffs(m) - 1;
            const char* hr = he + (size_t)S.g_slot[L] * 8;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(hr));
            if (S.g_size[L] > 8) asm volatile("prefetch.global.L2 [%0];" ::"l"(hr + 64));
        }
    }
    __syncwarp();  // rows (actions applied, masks, totals, zeroed histograms) are read by other lanes from here on

    // ---- combat, server.py:503-654: one lane per (match, fighting group) item; a round takes whole matches
    {
        const int nitems = __popc(fm);
        int incl = nitems;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += t;
        }
        const int pre = incl - nitems;
        const int total = __shfl_sync(0xFFFFFFFFu, incl, 31);
        int l_begin = 0;
        while (total && l_begin < 32) {
            const int base = __shfl_sync(0xFFFFFFFFu, pre, l_begin);
            const int l_end = __popc(__ballot_sync(0xFFFFFFFFu, incl - base <= 32)) & ~1;  // whole matches = lane pairs
            const int nround = __shfl_sync(0xFFFFFFFFu, incl, l_end - 1) - base;
            const int q = base + lane;
            int ow = 0;  // owner lane: the largest with pre[ow] <= q
#pragma unroll
            for (int step = 16; step; step >>= 1) {
                const int cand = ow + step;
                const int pc = __shfl_sync(0xFFFFFFFFu, pre, cand & 31);
                if (cand < 32 && pc <= q) ow = cand;
            }
            const int pw = __shfl_sync(0xFFFFFFFFu, pre, ow);
            const uint32_t fmo = __shfl_sync(0xFFFFFFFFu, fm, ow);
            const bool act = lane < nround;
            const int m = ow >> 1, side = ow & 1;
            uint32_t* Rm = wrow + (size_t)m * P;
            uint32_t* Xm = Rm + RW;
            const uint16_t* Xm16 = reinterpret_cast<const uint16_t*>(Xm);
            int L = 0, x = 1, tb = 0;
            uint32_t w0 = 0, w1 = 0;
            double hv[MAXSZ];
            double* hp = A.health;
            if (act) {
                const int gg = kth_set_bit(fmo, q - pw);
                L = side * G12 + gg;
                hp = A.health + (warp_env0 + m) * S.health_slots + S.g_slot[L];
                load_group<MAXSZ>(hp, S.g_size[L], hv);  // consumed after the draws
                w0 = Rm[2 * L];
                w1 = Rm[2 * L + 1];
                x = (int)(w0 & W0_LOC_MASK);
                const int cnt = __popc(w1 & 0xFFFFu);
                const uint32_t opp = Xm16[2 * (nn + x) + 1 - side], own = Xm16[2 * (nn + x) + side];
                const uint32_t n = opp & 0xFFu, hb = opp >> 8;  // opposing units at the node, their histogram base
                tb = (int)(own >> 8);
                // my group's range starts after the groups listed before it (arrival order, then gid:
                // node.groups[pid], :198,690-691); sibling counts are still pre-combat here
                const uint32_t key = (w1 >> 16) << 4 | (uint32_t)gg;
                for (uint32_t sm = (uint32_t)Xm16[2 * x + side] & ~(1u << gg); sm; sm &= sm - 1) {
                    const int g = __ffs(sm) - 1;
                    const uint32_t w1g = Rm[2 * (side * G12 + g) + 1];
                    if (((w1g >> 16) << 4 | (uint32_t)g) < key) tb += __popc(w1g & 0xFFFFu);
                }
                // draws, :549-566; 8 draws of 16 bits per Philox block (oracle/tape.py)
                const uint32_t dmg = S.g_damage[L];
                const uint32_t turn_m = Rm[kRecTurn] + 1u, ep_m = Rm[kRecEpisode];
                uint32_t* hw = Xm + 2 * nn + (1 - side) * S.pair_hwords;
                for (int b = 0; 8 * b < cnt; ++b) {
                    uint32_t r[4];
                    philox4x32_10(S.env_base + (uint32_t)(warp_env0 + m), turn_m,
                                  (uint32_t)x | (uint32_t)side << 8 | (uint32_t)gg << 16 | (uint32_t)b << 24, ep_m << 8, S.seed_lo, S.seed_hi, r);
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        if (8 * b + k < cnt) {
                            const uint32_t half = (k & 1) ? r[k >> 1] >> 16 : r[k >> 1] & 0xFFFFu;
                            const uint32_t idx = hb + ((half * n) >> 16);
                            if (sizeof(HistT) == 1) atomicAdd(&hw[idx >> 2], dmg << ((idx & 3u) * 8));
                            else atomicAdd(&hw[idx >> 1], dmg << ((idx & 1u) * 16));
                        }
                }
            }
            __syncwarp();
            // apply, :573-643: both sides drew on pre-combat counts (:572); one lane updates one whole group
            if (act) {
                const uint32_t nwd = Rm[kRecNode0 + x - 1];
                const int cb = (int)(int8_t)((nwd >> 16) & 0xFFu);
                const int bonus = (cb == side ? 1 : 0) + ((S.node_flags[x] >> 2) & 1);
                const int type = S.g_type[L];
                const double divisor = __dadd_rn(S.unit_armor[type], __dmul_rn((double)bonus, S.node_def[x]));
                const double* ltab = S.loss_tab + ((size_t)(type * nn + x) * 3 + bonus) * kLossD;
                const HistT* hist = reinterpret_cast<const HistT*>(Xm + 2 * nn + side * S.pair_hwords);
                int avg;
                const uint32_t alive = apply_group<MAXSZ, HistT>(hp, hv, S.g_size[L], w1 & 0xFFFFu, hist, tb, ltab, divisor, &avg);
                Rm[2 * L + 1] = (w1 & 0xFFFF0000u) | alive;  // alive == 0: destroyed, leaves the node list (:623-627)
                Rm[2 * L] = (w0 & ~(127u << W0_AVG_SHIFT)) | ((uint32_t)avg & 127u) << W0_AVG_SHIFT;
            }
            __syncwarp();
            l_begin = l_end;
        }
    }

    // ---- movement of MY groups (server.py:656-706) fused with my side's per-node sums:
    //   [0:10) units of all listed groups (:446-449), [10:24) count*control of non-moving groups (:718-724),
    //   [24:29) number of non-moving groups (:725-726); plus my unit points (:313-317)
    int sc[2] = {0, 0};  // my contributions to the two players' scores
    bool any_alive = false, basecap = false;
    if (valid) {
        uint32_t* acc = X + sd * nn;
        for (int i = 0; i < nn; ++i) acc[i] = 0;
        int pts_own = 0;
#pragma unroll 4
        for (int g = 0; g < G12; ++g) {
            const int L = sd * G12 + g;
            const uint2 w = *reinterpret_cast<const uint2*>(R + 2 * L);
            uint32_t w0 = w.x;
            const uint32_t alive = w.y & 0xFFFFu;
            if (alive) {  // destroyed groups are skipped, :663
                if (w0 & W0_READY) {
                    w0 = (w0 & ~W0_READY) | W0_MOVING;  // first turn only flips ready -> moving (:664-667)
                } else if (w0 & W0_MOVING) {
                    const int dist = (int)((w0 >> W0_DIST_SHIFT) & 0xFFu) - (int)S.g_speed[L];  // :671
                    if (dist <= 0) {  // arrived: appended to the destination's list (:678-695)
                        w0 = (w0 & (127u << W0_AVG_SHIFT)) | ((w0 >> W0_DEST_SHIFT) & 0x3Fu);
                        R[2 * L + 1] = alive | turn << 16;
                    } else {
                        w0 = (w0 & ~(0xFFu << W0_DIST_SHIFT)) | (uint32_t)dist << W0_DIST_SHIFT;
                    }
                }
                R[2 * L] = w0;
                const uint32_t cnt = __popc(alive);
                uint32_t v = cnt;
                if (!(w0 & W0_MOVING)) v |= (cnt * S.g_control[L]) << 10 | 1u << 24;
                acc[w0 & W0_LOC_MASK] += v;
                pts_own += (int)cnt * (int)S.g_cost[L];
                any_alive = true;
            }
        }
        sc[sd] = pts_own;
    }
    __syncwarp();  // the partner's sums

    // ---- capture (server.py:708-767; current_turn > 0 here) and node scoring (:298-310): nodes split by parity
    if (valid) {
        for (int n = 1 + sd; n <= n_nodes; n += 2) {
            uint32_t nw = R[kRecNode0 + n - 1];
            int cs = (int)(int16_t)(nw & 0xFFFFu), cb = (int)(int8_t)((nw >> 16) & 0xFFu);
            const uint32_t a0 = X[n], a1 = X[nn + n];
            const bool c0 = (a0 >> 24) != 0, c1 = (a1 >> 24) != 0;
            const int cp = S.node_cp[n];
            if (c0 != c1) {  // exactly one controller (:729)
                const int pid = c1 ? 1 : 0;
                if (abs(cs) < cp || pid != cb) {  // :731-732
                    const int pts = (int)(((pid ? a1 : a0) >> 10) & 0x3FFFu), pxer = pid ? -1 : 1;
                    const bool old_sign = cs < 0;  // :747-750, zero counts as player 0's sign
                    cs += pts * pxer;
                    const bool neutralize = old_sign != (cs < 0);
                    if (abs(cs) >= cp) {  // :763-765
                        cs = cp * pxer;
                        cb = pid;
                    }
                    if (cb != -1 && neutralize) cb = -1;  // :766-767
                    nw = ((uint32_t)cs & 0xFFFFu) | ((uint32_t)cb & 0xFFu) << 16;
                    R[kRecNode0 + n - 1] = nw;
                }
            }
            const int ts = S.node_team_start[n];
            if (ts != -1 && cb != -1 && cb != ts) {
                basecap = true;
                sc[cb] += S.capture_bonus;
            }
            if (cs != 0) sc[cs > 0 ? 0 : 1] += abs(cs) == cp ? 2 * cp : abs(cs);
        }
    }
    // both lanes of a pair now combine their halves
    const int s0 = sc[0] + __shfl_xor_sync(0xFFFFFFFFu, sc[0], 1);
    const int s1 = sc[1] + __shfl_xor_sync(0xFFFFFFFFu, sc[1], 1);
    const int partner_alive = __shfl_xor_sync(0xFFFFFFFFu, (int)any_alive, 1);  // (no short-circuit around a shuffle)
    const int partner_cap = __shfl_xor_sync(0xFFFFFFFFu, (int)basecap, 1);
    any_alive = any_alive || partner_alive;
    basecap = basecap || partner_cap;
    int status = EVG_STATUS_IN_PROGRESS;  // server.py:321-328, in that priority
    if ((int)turn >= S.turn_limit) status = EVG_STATUS_TIME_EXPIRED;
    else if (!any_alive) status = EVG_STATUS_ANNIHILATION;
    else if (basecap) status = EVG_STATUS_BASE_CAPTURE;
    const bool done = valid && status != 0;

    // ---- reward / done, env.py:37-60: each lane its own player's reward (float32 division == float32(float64 q))
    if (valid) {
        const int mine = sd ? s1 : s0, other = sd ? s0 : s1;
        float r;
        if (done) r = mine == other ? 0.f : (mine > other ? 1.f : (sd ? -1.f : 0.f));
        else r = __fdiv_rn((float)mine, S.max_score_f);
        A.reward[env * 2 + sd] = r;
        if (A.scores) A.scores[env * 2 + sd] = mine;
        if (sd == 0) {
            A.done[env] = done ? 1 : 0;
            if (A.status) A.status[env] = (uint8_t)status;
        }
    }
    const bool reset_now = done && S.auto_reset != EVG_AUTORESET_OFF;
    // ---- episode end: statistics (player-0 lanes), aggregated over the warp before the global counters
    if (__any_sync(0xFFFFFFFFu, reset_now)) {
        const unsigned e = (reset_now && sd == 0) ? 1u : 0u;
        const unsigned v[ST_COUNT] = {e, e && s0 > s1, e && s1 > s0, e && s0 == s1, e ? turn : 0u, e ? (unsigned)s0 : 0u,
                                      e ? (unsigned)s1 : 0u, e && status == 0, e && status == 1, e && status == 2, e && status == 3};
#pragma unroll
        for (int k = 0; k < ST_COUNT; ++k) {
            const unsigned sum = __reduce_add_sync(0xFFFFFFFFu, v[k]);
            if (lane == 0 && sum) atomicAdd(&A.stats[k], (unsigned long long)sum);
        }
    }
    if (reset_now && S.auto_reset == EVG_AUTORESET_NEXT) {  // show the NEW match's first observation
        reset_side(S, R, A.health + env * S.health_slots + sd * hslots_side, n_nodes, sd);
        turn = 0;
        episode += 1;
        uint32_t* acc = X + sd * nn;
        for (int i = 0; i < nn; ++i) acc[i] = 0;
        for (int g = 0; g < G12; ++g) {
            const int L = sd * G12 + g;
            const uint32_t cnt = __popc(R[2 * L + 1] & 0xFFFFu);
            acc[R[2 * L] & W0_LOC_MASK] += cnt | (cnt * S.g_control[L]) << 10 | 1u << 24;
        }
    }
    __syncwarp();  // node words and per-node sums of both lanes are final

    // ---- observations: board_state (server.py:382-455) + player_state (:457-501) + concat (env.py:158-171).
    // Each lane packs ITS player's 105 values, 16 at a time, into its 64-byte window of the row; the warp
    // streams the 32 windows out as contiguous runs (two windows = 128 bytes per store instruction).
    {
        const int stage_off = RW + 2 * nn;
        float* stage = reinterpret_cast<float*>(X + 2 * nn + 16 * sd);
        float* obs_base = A.obs + warp_env0 * 2 * OL;
        auto value = [&](int i) -> float {  // i = index into my player's OL values
            if (i == 0) return (float)turn;
            if (i < 1 + 4 * n_nodes) {
                const int k = (i - 1) >> 2, j = (i - 1) & 3;
                const int x = sd ? (int)S.p1_map[k + 1] : k + 1;  // server.py:437-439
                if (j == 0) return (float)(S.node_flags[x] & 1u);
                if (j == 1) return (float)((S.node_flags[x] >> 1) & 1u);
                if (j == 2) return (float)(int)(int16_t)(R[kRecNode0 + x - 1] & 0xFFFFu);  // raw sign for both viewers
                return (float)(X[(sd ? 0 : nn) + x] & 1023u);                                 // opposing listed units
            }
            const int q = i - 1 - 4 * n_nodes, g = q / 5, j = q - 5 * g;
            const int L = sd * G12 + g;
            const uint32_t w0 = R[2 * L];
            if (j == 0) return (float)(sd ? (uint32_t)S.p1_map[w0 & W0_LOC_MASK] : (w0 & W0_LOC_MASK));
            if (j == 1) return (float)S.g_type[L];
            if (j == 2) return (float)((w0 >> W0_AVG_SHIFT) & 127u);
            if (j == 3) return (float)((w0 >> 21) & 1u);
            return (float)__popc(R[2 * L + 1] & 0xFFFFu);
        };
        const int hw = lane >> 4, c16 = lane & 15;
        const int nchunks = (OL + 15) / 16;
#pragma unroll
        for (int c = 0; c < (NODES ? (1 + 4 * NODES + 60 + 15) / 16 : nchunks); ++c) {
            if (valid) {
                float vals[16];  // all reads first (they can be merged and overlapped), then the window stores
#pragma unroll
                for (int k = 0; k < 16; ++k) vals[k] = 16 * c + k < OL ? value(16 * c + k) : 0.f;
#pragma unroll
                for (int k = 0; k < 16; k += 2) *reinterpret_cast<float2*>(stage + k) = make_float2(vals[k], vals[k + 1]);
            }
            __syncwarp();
            if (16 * c + c16 < OL) {
#pragma unroll
                for (int it = 0; it < 16; ++it) {
                    const int wdw = 2 * it + hw;  // window = (match, player)
                    if ((wdw >> 1) < nvalid) {
                        const float v = *reinterpret_cast<const float*>(wrow + (size_t)(wdw >> 1) * P + stage_off + 16 * (wdw & 1) + c16);
                        __stcs(obs_base + (size_t)wdw * OL + 16 * c + c16, v);
                    }
                }
            }
            __syncwarp();
        }
    }
    if (reset_now && S.auto_reset == EVG_AUTORESET_TERMINAL) {
        reset_side(S, R, A.health + env * S.health_slots + sd * hslots_side, n_nodes, sd);
        turn = 0;
        episode += 1;
    }
    if (valid && sd == 0) {
        R[kRecTurn] = turn;
        R[kRecEpisode] = episode;
    }
    __syncwarp();

    // ---- cooperative, coalesced store of the records
    {
        const int q4 = RW / 4;
        uint4* g4 = reinterpret_cast<uint4*>(A.records) + warp_env0 * q4;
        const int total = nvalid * q4;
#pragma unroll 8
        for (int f = lane; f < total; f += 32) {
            const int m = NODES ? f >> 4 : f / q4, q = NODES ? f & 15 : f % q4;
            const uint2* s = reinterpret_cast<const uint2*>(wrow + (size_t)m * P + 4 * q);
            const uint2 a = s[0], b = s[1];
            g4[f] = make_uint4(a.x, a.y, b.x, b.y);
        }
    }
    __syncwarp();
    }  // batch loop
}

enum Variant { V_FAST = 0, V_GENERIC8, V_GENERIC16 };

Variant pick(const Tables& t)
{
    if (t.n_nodes == 11 && t.max_group_size <= 12 && !t.tpm_hist16 && t.pair_pitch == 138) return V_FAST;
    return t.tpm_hist16 ? V_GENERIC16 : V_GENERIC8;
}

}  // namespace

cudaError_t pair_prepare(const Tables& t, size_t* smem_out, int* blocks_per_sm)
{
    const size_t smem = (size_t)t.sm_tables_bytes + 128 + (size_t)kPairMatches * t.pair_pitch * 4;
    *smem_out = smem;
    cudaError_t e;
    int limit = 0;
    if ((e = optin_smem_limit(smem, &limit)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(evg_step_pair_kernel<11, 12, uint8_t, 138>, cudaFuncAttributeMaxDynamicSharedMemorySize, limit)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(evg_step_pair_kernel<0, 16, uint8_t, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, limit)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(evg_step_pair_kernel<0, 16, uint16_t, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, limit)) != cudaSuccess) return e;
    switch (pick(t)) {
        case V_FAST: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, evg_step_pair_kernel<11, 12, uint8_t, 138>, kPairThreads, smem); break;
        case V_GENERIC8: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, evg_step_pair_kernel<0, 16, uint8_t, 0>, kPairThreads, smem); break;
        default: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, evg_step_pair_kernel<0, 16, uint16_t, 0>, kPairThreads, smem); break;
    }
    return e;
}

cudaError_t launch_step_pair(const Tables& t, const StepArgs& a, size_t smem, int max_grid, cudaStream_t stream)
{
    const int64_t nb = (a.n_envs + kPairMatches - 1) / kPairMatches;
    const unsigned grid = (unsigned)(nb < max_grid ? nb : max_grid);
    switch (pick(t)) {
        case V_FAST: evg_step_pair_kernel<11, 12, uint8_t, 138><<<grid, kPairThreads, smem, stream>>>(t, a); break;
        case V_GENERIC8: evg_step_pair_kernel<0, 16, uint8_t, 0><<<grid, kPairThreads, smem, stream>>>(t, a); break;
        default: evg_step_pair_kernel<0, 16, uint16_t, 0><<<grid, kPairThreads, smem, stream>>>(t, a); break;
    }
    return cudaGetLastError();
}

}  // namespace evg
