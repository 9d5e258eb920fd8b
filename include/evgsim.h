/*
 * evgsim.h — C ABI of libevgsim: the batched, GPU-resident Everglades turn step.
 *
 * This is the drop-in boundary for ONE path of jlehett/everglades-ai-wargame:
 *     EvergladesEnv.reset / EvergladesEnv.step
 *         gym-everglades/gym_everglades/envs/everglades_env.py:32-116   ("env.py")
 *     -> EvergladesGame.game_init / game_turn / board_state / player_state
 *         everglades-server/everglades_server/server.py:133-501          ("server.py")
 * for N matches in lockstep.  The reference has no FFI (it is pure Python); the binding a
 * maintainer would add is a ctypes stub — see INTEGRATION.md.
 *
 * Rules of the boundary
 *   - plain C types only; no torch / CUDA runtime types in any signature (`stream` is a
 *     cudaStream_t passed as void*, 0 = the legacy default stream);
 *   - the library NEVER allocates device memory: the caller asks evg_layout() for sizes,
 *     allocates (torch.empty / cudaMalloc) and binds raw device pointers with evg_bind();
 *   - every call is asynchronous on `stream`; nothing synchronises unless stated;
 *   - every entry point returns 0 on success or a negative EVG_E_* code; evg_last_error()
 *     returns a thread-local human-readable message for the last failure.
 *   - there is NO CPU fallback: without a CUDA device evg_create() fails with EVG_E_CUDA.
 */
#ifndef EVGSIM_H
#define EVGSIM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EVG_ABI_VERSION 1

/* compile-time maxima of the state layout */
#define EVG_MAX_NODES 32       /* node IDs are 1..n_nodes (DemoMap: 11) */
#define EVG_MAX_UNIT_TYPES 8   /* UnitDefinitions.json entries (reference: 3) */
#define EVG_NUM_GROUPS 12      /* groups per player, env.py:19 */
#define EVG_MAX_GROUP_UNITS 16 /* unit slots per group (reference loadout: 8, last group 12) */
#define EVG_MAX_ACTIONS 7      /* rows of an action array that count, server.py:227 */
#define EVG_NUM_PLAYERS 2

/* error codes */
#define EVG_OK 0
#define EVG_E_ARG -1    /* null pointer / out-of-range argument */
#define EVG_E_CONFIG -2 /* EvgConfig fails validation (see evg_last_error) */
#define EVG_E_CUDA -3   /* CUDA runtime error (no device, launch failure, ...) */
#define EVG_E_STATE -4  /* call sequence error, e.g. step before bind */

/* game status, server.py:284-289 */
#define EVG_STATUS_IN_PROGRESS 0
#define EVG_STATUS_TIME_EXPIRED 1
#define EVG_STATUS_BASE_CAPTURE 2
#define EVG_STATUS_ANNIHILATION 3

/* what evg_step does with a match whose status != 0 */
#define EVG_AUTORESET_OFF 0       /* keep stepping it, like the reference server does */
#define EVG_AUTORESET_TERMINAL 1  /* reset in place; obs of that step = terminal obs (reference-identical) */
#define EVG_AUTORESET_NEXT 2      /* reset in place; obs of that step = first obs of the new match */

/*
 * Static game description: DemoMap.json + UnitDefinitions.json + GameSetup.json semantics
 * (server.py:40-131 board_init/unitTypes_init; env.py:15-22,145-156 constants and loadout).
 * All per-node arrays are indexed by node ID (entry 0 unused).
 */
typedef struct EvgConfig {
    int32_t abi_version; /* = EVG_ABI_VERSION */
    int32_t n_nodes;
    int32_t n_unit_types;
    int32_t turn_limit;    /* GameSetup.json TurnLimit; hard-coded 150 at server.py:321 */
    int32_t capture_bonus; /* GameSetup.json CaptureBonus; hard-coded 1000 at server.py:304 */
    int32_t max_score;     /* reward normaliser MAX_SCORE = 3700, env.py:11 */
    int32_t auto_reset;    /* EVG_AUTORESET_* */
    int32_t reserved0;

    int32_t node_control_points[EVG_MAX_NODES + 1]; /* "ControlPoints" */
    double node_defense[EVG_MAX_NODES + 1];         /* "StructureDefense" */
    int8_t node_team_start[EVG_MAX_NODES + 1];      /* "TeamStart": -1, 0 or 1 */
    uint8_t node_has_defense[EVG_MAX_NODES + 1];    /* 'DEFENSE' in Resource: obs flag, server.py:442 */
    uint8_t node_has_observe[EVG_MAX_NODES + 1];    /* 'OBSERVE' in Resource: obs flag, server.py:443 */
    uint8_t node_has_defend[EVG_MAX_NODES + 1];     /* 'DEFEND' in Resource: combat bonus, server.py:595
                                                       (never true on DemoMap, whose maps say "DEFENSE") */
    /* edge_distance[a][b] = "Distance" of the FIRST connection of node a whose ConnectedID is b
       (server.py:246-250), 0 = not connected */
    uint8_t edge_distance[EVG_MAX_NODES + 1][EVG_MAX_NODES + 1];
    /* player 1's view of node ids (server.py:89); must be an involution with map[0] = 0 */
    uint8_t p1_node_map[EVG_MAX_NODES + 1];

    double unit_armor[EVG_MAX_UNIT_TYPES];    /* "Health" — used as armour, server.py:592 */
    int32_t unit_damage[EVG_MAX_UNIT_TYPES];  /* "Damage" */
    int32_t unit_speed[EVG_MAX_UNIT_TYPES];   /* "Speed" */
    int32_t unit_control[EVG_MAX_UNIT_TYPES]; /* "Control" */
    int32_t unit_cost[EVG_MAX_UNIT_TYPES];    /* "Cost" */

    /* loadout, env.py:145-156: type id = index in UnitDefinitions.json (server.py:128-129) */
    uint8_t group_type[EVG_NUM_PLAYERS][EVG_NUM_GROUPS];
    uint8_t group_size[EVG_NUM_PLAYERS][EVG_NUM_GROUPS]; /* 1..EVG_MAX_GROUP_UNITS */
} EvgConfig;

/* One group, as the reference keeps it (definitions.py:35-47 EvgGroup + :58-64 EvgUnit). */
typedef struct EvgGroupState {
    int16_t location;           /* node ID */
    int16_t travel_destination; /* node ID or -1 */
    int16_t distance_remaining;
    uint8_t ready;
    uint8_t moving;
    uint8_t destroyed;
    uint8_t count;       /* alive units */
    int32_t arrival;     /* list-order stamp: groups of one player at one node are listed by
                            ascending arrival (node.groups[pid], server.py:198,690-691);
                            canonical value: gid at reset, turn*16+gid on arrival */
    int32_t avg_health;  /* int(sum(health)/alive) as player_state reports it (server.py:491) */
} EvgGroupState;

/* One match, array-of-structs view used by evg_export_state / evg_import_state (parity tests,
 * checkpointing).  The resident device layout is different (see evg_layout). */
typedef struct EvgEnvState {
    int32_t turn;    /* current_turn, server.py:140,214 */
    int32_t episode; /* how many matches this slot has finished or been reset out of; keys the tape
                        so that successive matches of one slot differ (0 after evg_reset(NULL)) */
    int16_t control_state[EVG_MAX_NODES + 1]; /* node.controlState, definitions.py:17 */
    int8_t controlled_by[EVG_MAX_NODES + 1];  /* node.controlledBy, definitions.py:16 */
    int8_t pad1[5];
    EvgGroupState groups[EVG_NUM_PLAYERS][EVG_NUM_GROUPS];
    double health[EVG_NUM_PLAYERS][EVG_NUM_GROUPS][EVG_MAX_GROUP_UNITS]; /* unitHealth, definitions.py:62 */
} EvgEnvState;

/* Sizes of the device arrays the caller must allocate and bind (bytes, for n_envs matches). */
typedef struct EvgLayout {
    int64_t n_envs;
    int32_t obs_len;       /* per player: 1 + 4*n_nodes + 5*12 (105 on DemoMap), env.py:165-167 */
    int32_t record_bytes;  /* per match: packed group/node/turn record */
    int32_t health_slots;  /* per match: fp64 unit-health slots (both players, padded) */
    int32_t action_bytes;  /* per match: 2*7*2 int8 */
    int64_t records_bytes; /* = n_envs * record_bytes           (bind slot EVG_BIND_RECORDS) */
    int64_t health_bytes;  /* = n_envs * health_slots * 8       (bind slot EVG_BIND_HEALTH)  */
    int64_t stats_bytes;   /* episode statistics accumulators   (bind slot EVG_BIND_STATS)   */
    int64_t tables_bytes;  /* derived lookup tables, filled by evg_bind (bind slot EVG_BIND_TABLES) */
    int64_t agents_bytes;  /* scripted-agent state, 16 B per match (bind slot EVG_BIND_AGENTS)     */
} EvgLayout;

#define EVG_BIND_RECORDS 0
#define EVG_BIND_HEALTH 1
#define EVG_BIND_STATS 2
#define EVG_BIND_TABLES 3
#define EVG_BIND_AGENTS 4
#define EVG_BIND_COUNT 5

/* Episode statistics accumulated on the device by evg_step (matches that ended). */
typedef struct EvgEpisodeStats {
    int64_t episodes;
    int64_t wins[2]; /* by final score, the test callers use (evaluate.py:155-160) */
    int64_t ties;
    int64_t total_turns;     /* sum of episode lengths */
    int64_t total_score[2];  /* sum of final scores */
    int64_t status_count[4]; /* histogram of EVG_STATUS_* at episode end */
    int64_t env_turns;       /* all match-turns stepped since creation / last clear */
    int64_t fought_unit_slots; /* unit slots of the groups that took part in combat, summed over all match-turns
                                  stepped (finished matches or not): 16 B of fp64 health traffic each, the variable
                                  term of the step's algorithmic bytes (server.py:516-566) */
} EvgEpisodeStats;

typedef struct EvgSim EvgSim; /* opaque */

/* Fill `cfg` with the reference's DemoMap.json / UnitDefinitions.json / GameSetup.json values
 * and the env.py:145-156 loadout (useful when no JSON parser is at hand). */
int evg_default_config(EvgConfig* cfg);

/* Validate `cfg` and create a simulator for `n_envs` matches on CUDA device `device`.
 * Match i has global id env_id_offset + i; the combat tape is keyed on (seed, global id,
 * turn, node, side, group, unit) so trajectories do not depend on how matches are sharded.
 * Replaces: EvergladesEnv.__init__ (env.py:15-30) + EvergladesGame.__init__ (server.py:14-38). */
int evg_create(const EvgConfig* cfg, int64_t n_envs, uint64_t seed, int64_t env_id_offset, int device,
               EvgSim** out);
int evg_destroy(EvgSim* sim);

int evg_layout(const EvgSim* sim, EvgLayout* out);
/* Bind caller-owned device arrays (index = EVG_BIND_*). Must precede reset/step.  Uploads the derived
 * tables into slot EVG_BIND_TABLES (synchronous copy). */
int evg_bind(EvgSim* sim, void* const* device_ptrs, int32_t n_ptrs);

/* Put matches into the post-game_init state (server.py:133-209, env.py:75-116) and write their
 * observations.  d_mask: n_envs bytes, non-zero = reset this match; NULL = all.
 * d_obs: float32 [n_envs][2][obs_len] or NULL.  Also zeroes the statistics when d_mask is NULL. */
int evg_reset(EvgSim* sim, const uint8_t* d_mask, float* d_obs, void* stream);

/* One game turn for every match (env.py:32-73 step -> server.py:211-279 game_turn, 281-348
 * game_end, 382-501 observations).
 *   d_actions: int8 [n_envs][2][7][2] = (group id, node id in the acting player's numbering);
 *              rows with group/node out of range are ignored (the reference raises IndexError).
 *   d_obs:     float32 [n_envs][2][obs_len]   (reference: float64 arrays of integers)
 *   d_reward:  float32 [n_envs][2]            (env.py:37-60)
 *   d_done:    uint8   [n_envs]               (status != 0)
 *   d_status / d_scores: optional (may be NULL): uint8 [n_envs], int32 [n_envs][2]. */
int evg_step(EvgSim* sim, const int8_t* d_actions, float* d_obs, float* d_reward, uint8_t* d_done,
             uint8_t* d_status, int32_t* d_scores, void* stream);

/*
 * Observation formats.  EVG_OBS_F32 is the reference's vector (env.py:158-171) as float32.  The other two carry the SAME
 * information losslessly in fewer bytes, for consumers behind a narrow link (the host over PCIe: evg_step_host_fmt):
 *
 *   EVG_OBS_I16   int16 [n_envs][2][obs_len], same layout as the float32 vector (every entry is an integer in
 *                 [-32768, 32767]: turn <= 65535 is checked at create time against this format on use).
 *   EVG_OBS_WIRE  one packed row per match, evg_obs_row_bytes() bytes = round_up(4 + 4*n_nodes + 72 + 8, 16)
 *                 (128 on DemoMap: one cache line), little endian:
 *        [0:2)  turn uint16 (obs[0], server.py:432)    [2] done uint8    [3] status uint8 (EVG_STATUS_*)
 *        [4 + 4*(x-1) ...) for node id x = 1..n_nodes: controlState int16 (raw sign, server.py:444), units of player 0's
 *                 groups listed at the node uint8, units of player 1's groups uint8 (server.py:446-449; a viewer's
 *                 "opponent units" entry is the other player's byte)
 *        [G = 4 + 4*n_nodes + 3*(12*p + g) ...) group g of player p: byte 0 = location in REAL node numbering [0:6) |
 *                 moving << 6 (server.py:477,489), byte 1 = avg health (server.py:491), byte 2 = units alive (:480)
 *        [G + 72 : G + 80)  the step's rewards, float32[2] (env.py:37-60)
 *        zero padding to the row size.
 *   What the float32 vector holds besides is static: the nodes' DEFENSE/OBSERVE flags, the groups' unit types, and
 *   player 1's node numbering (server.py:437-439, 476) — a function of EvgConfig alone.  evgsim.wire.expand() /
 *   INTEGRATION.md give the expansion; tests/test_gpu_wire.py checks expand(wire) == float32 observations bit for bit.
 */
#define EVG_OBS_F32 0
#define EVG_OBS_I16 1
#define EVG_OBS_WIRE 2
#define EVG_WIRE_NODE0 4 /* byte offset of node 1's entry in a wire row */

/* where a player's action rows come from in evg_step_agents */
#define EVG_AGENT_EXTERNAL 0 /* the caller's rows in d_actions */
#define EVG_AGENT_RANDOM 1   /* on-device random_actions agent (agents/State_Machine/random_actions.py:38-46) */
#define EVG_AGENT_BASE_RUSH 2 /* base_rushV1 (agents/State_Machine/base_rush_v1.py:62-111); keeps its counters per match */
#define EVG_AGENT_SWARM 3    /* SwarmAgent (agents/State_Machine/swarm_agent.py:79-102); keeps its attack list per match */

/* evg_step with scripted opponents fused into the step kernel: rows of players whose agent is not
 * EVG_AGENT_EXTERNAL are generated on the device (and written to d_actions if it is non-NULL); rows of
 * EXTERNAL players are read from d_actions (then it must be non-NULL).  With both players scripted a
 * whole self-play turn is ONE kernel launch and no action buffer is touched. */
int evg_step_agents(EvgSim* sim, int32_t agent_p0, int32_t agent_p1, int8_t* d_actions, float* d_obs, float* d_reward,
                    uint8_t* d_done, uint8_t* d_status, int32_t* d_scores, void* stream);

/* `n_turns` turns of fully scripted self-play (both agents != EVG_AGENT_EXTERNAL) with the results of evg_step_agents
 * called n_turns times: the output arrays hold the last turn's values, statistics accumulate, d_actions (required) is
 * scratch for the generated rows.  ALL the turns run in ONE launch: small batches (evg_step_kernel_kind() == 0) on the
 * multi-turn warp-per-match kernel (a warp keeps its match for the whole rollout), larger ones on the thread-per-match
 * kernel, whose CTAs keep each batch in shared memory for all n_turns — a turn costs neither a launch nor an action-buffer
 * pass, and only the last one writes observations and records.  (Only EVG_AGENT_RANDOM on maps of more than 15 nodes
 * runs turn by turn: agent kernel + step.) */
int evg_rollout(EvgSim* sim, int32_t agent_p0, int32_t agent_p1, int32_t n_turns, int8_t* d_actions, float* d_obs, float* d_reward,
                uint8_t* d_done, uint8_t* d_status, int32_t* d_scores, void* stream);

/* Same turn through HOST buffers: H2D of the actions, the step, D2H of obs/reward/done, ordered after
 * what `stream` holds and complete when `stream` is (pinned host memory makes them truly asynchronous).
 * The device staging arrays are the caller's (same shapes as evg_step).  From 65,536 matches on, the
 * thread-per-match kernel runs the batch as EVG_HOST_CHUNKS (environment, default 16) sub-range launches on
 * two streams the library creates for this, so that the PCIe-bound D2H of one chunk hides the H2D and
 * the kernel of the next; results are those of evg_step. */
int evg_step_host(EvgSim* sim, const int8_t* h_actions, float* h_obs, float* h_reward, uint8_t* h_done,
                  int8_t* d_actions, float* d_obs, float* d_reward, uint8_t* d_done, void* stream);

/* Bytes per match of an observation format (0 for an unknown format or NULL sim). */
int evg_obs_row_bytes(const EvgSim* sim, int32_t format);

/* evg_reset / evg_step / evg_step_host with the observation output in `format` (EVG_OBS_*); with EVG_OBS_F32 they ARE
 * those calls.  d_rows / h_rows: n_envs rows of evg_obs_row_bytes(sim, format).  EVG_OBS_WIRE rows are written by the
 * step kernel itself (no float32 observations are produced at all); EVG_OBS_I16 needs a caller-provided float32
 * scratch d_obs_f32 [n_envs][2][obs_len] that the step writes and a conversion kernel narrows.  In evg_step_host_fmt
 * h_reward / h_done may be NULL with EVG_OBS_WIRE (the row carries them). */
int evg_reset_fmt(EvgSim* sim, int32_t format, const uint8_t* d_mask, void* d_rows, float* d_obs_f32, void* stream);
int evg_step_fmt(EvgSim* sim, int32_t format, const int8_t* d_actions, void* d_rows, float* d_obs_f32, float* d_reward,
                 uint8_t* d_done, uint8_t* d_status, int32_t* d_scores, void* stream);
int evg_step_host_fmt(EvgSim* sim, int32_t format, const int8_t* h_actions, void* h_rows, float* h_reward, uint8_t* h_done,
                      int8_t* d_actions, void* d_rows, float* d_obs_f32, float* d_reward, uint8_t* d_done, void* stream);

/* Array-of-structs snapshot <-> resident layout, matches [first, first+count).  d_states is a
 * DEVICE array of EvgEnvState (the caller copies it to/from the host). */
int evg_export_state(EvgSim* sim, int64_t first, int64_t count, EvgEnvState* d_states, void* stream);
/* Import validates what it is given: a location outside 1..n_nodes, a travel_destination above n_nodes, a
 * distance_remaining outside 0..255, |control_state| above the node's ControlPoints or a controlled_by outside -1..1 is
 * forced into range (these fields index tables in the step kernels) and the call returns EVG_E_ARG after importing the
 * rest.  Synchronises `stream`. */
int evg_import_state(EvgSim* sim, int64_t first, int64_t count, const EvgEnvState* d_states, void* stream);

/* Copy the statistics accumulators to the host (synchronises `stream`). */
int evg_episode_stats(EvgSim* sim, EvgEpisodeStats* host_out, void* stream);

/* On-device scripted opponents (agents/State_Machine, Python files), writing int8 [n_envs][2][7][2]
 * action rows for `player` (0, 1, or -1 = both).  See DESIGN.md §7 for their tape. */
int evg_agent_random(EvgSim* sim, int8_t* d_actions, int32_t player, void* stream);

/* Action rows of the scripted agents for both players (EVG_AGENT_EXTERNAL players are left untouched) into
 * int8 [n_envs][2][7][2].  base_rushV1 / SwarmAgent carry per-match state across turns and matches (like the
 * reference's agent objects); evg_reset(sim, NULL, ...) makes all of them fresh agents again. */
int evg_agents(EvgSim* sim, int32_t agent_p0, int32_t agent_p1, int8_t* d_actions, void* stream);

/* Policy-in-the-loop glue: decode network outputs into action rows on the device (player: 0, 1, or -1 = both;
 * inputs are then [n_envs][2][...], else [n_envs][...]).
 *   evg_decode_dqn:     float32 Q-values [..][12 * num_cols] -> rows, exactly as DQNAgent.filter_actions does
 *                       (agents/DQN/DQNAgent.py:161-197; greedy insertion, node = 0-based column index).
 *   evg_decode_indices: int64 flat indices [..][7] -> rows (idx / div, idx % mod); PPOAgent.get_action uses
 *                       div 12, mod 11 (agents/PPO/PPOAgent.py:122-127), DQN's replay encoding div 11, mod 11. */
int evg_decode_dqn(EvgSim* sim, const float* d_q, int32_t num_cols, int32_t player, int8_t* d_actions, void* stream);
/* the same with d_q optionally transposed: [12 * num_cols][rows], rows = n_envs * (player < 0 ? 2 : 1) */
int evg_decode_dqn_layout(EvgSim* sim, const float* d_q, int32_t num_cols, int32_t player, int32_t q_transposed, int8_t* d_actions, void* stream);
int evg_decode_indices(EvgSim* sim, const int64_t* d_idx, int32_t div, int32_t mod, int32_t player, int8_t* d_actions,
                       void* stream);

/* Policy-in-the-loop forward, fused on the tensor cores (tcgen05 + TMEM; csrc/evg_policy_mlp.cu):
 *     d_q[rows][out_dim] = Linear(hidden, out_dim)(relu(Linear(obs_len, hidden)(d_obs[rows][obs_len])))
 * — the reference's DQN network, agents/DQN/QNetwork.py:37,42 (105 -> 528 -> 132) — for the observation rows the step
 * has just written (rows = n_envs * 2 for both players), bf16 operands with fp32 accumulation; the hidden activations stay
 * on the SM.  Weights are passed as IMAGES in the kernel's shared-memory operand layout (bf16, K-major, 128-byte swizzle),
 * cut into chunks of EVG_MLP_CHUNK hidden units, with the biases inside (homogeneous coordinates): input feature obs_len
 * is the constant 1, hidden unit `hidden` is wired to relu(1 * 1) = 1.  With W1' = [W1 | b1] plus that unit's row, and
 * W2' = [W2 | b2], for chunk c (ceil((hidden + 1) / CHUNK) chunks):
 *     d_w1_img + c * (EVG_MLP_IN_PAD / 64) * EVG_MLP_CHUNK * 128 bytes:  W1'[c*CHUNK + n][k], n < CHUNK, k < IN_PAD
 *     d_w2_img + c * (EVG_MLP_CHUNK / 64) * EVG_MLP_OUT_PAD * 128 bytes: W2'[o][c*CHUNK + k], o < OUT_PAD, k < CHUNK
 * element (row r, column k) of a [R x K] block at byte (k/64)*R*128 + r*128 + ((((k%64)/8) ^ (r%8)) * 16) + (k%8)*2, zero
 * padded.  evgsim.policy.pack_mlp builds them from a torch module.  q_transposed != 0 writes d_q as [out_dim][rows]
 * instead (each output's values for all rows contiguous): the layout evg_decode_dqn_layout reads coalesced. */
#define EVG_MLP_IN_PAD 128
#define EVG_MLP_CHUNK 192
#define EVG_MLP_OUT_PAD 144
int evg_policy_mlp(EvgSim* sim, const float* d_obs, int64_t rows, const void* d_w1_img, const void* d_w2_img, int32_t hidden, int32_t out_dim,
                   float* d_q, int32_t q_transposed, void* stream);

/* Reward shaping of the reference's training scripts (utils/reward_shaping.py:17-56) on the step's outputs:
 * d_out float32 [n_envs][2].  turnNum (steps played before this one) is read from the observation's turn
 * field, so use it with EVG_AUTORESET_OFF or _TERMINAL (the observation of a finished match is the terminal one). */
#define EVG_SHAPE_NORMALIZED_SCORE 0
#define EVG_SHAPE_BASIC 1
#define EVG_SHAPE_PENALIZE_LONG 2
#define EVG_SHAPE_SHORT_GAMES 3
int evg_shape_reward(EvgSim* sim, int32_t mode, const float* d_reward, const uint8_t* d_done, const float* d_obs, float* d_out,
                     void* stream);

/* Which step kernel evg_create() selected: 0 = a warp per match (small batches), 1 = a thread per match
 * (DESIGN.md section 4).  Scripted agents are fused into kernel 1 only; with kernel 0 evg_step_agents() needs
 * a non-NULL d_actions to pass the generated rows through.  -1 if sim is NULL. */
int evg_step_kernel_kind(const EvgSim* sim);

/* Number of kernels this library has launched since creation (bench.py's gpu_launches). */
int64_t evg_launch_count(const EvgSim* sim);

const char* evg_last_error(void);
int evg_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* EVGSIM_H */
