// Fused non-physics MDP step of AAURoverEnv-v0: two launches around the (external) physics state.
//
//   mdp_pre_step_kernel   one thread per env: action manager shift, AckermannAction2 kinematics, counters,
//                         the 4 terminations, the 7 weighted rewards + episodic sums, reset flags.
//   mdp_post_step_kernel  one thread per env: rank of each reset env (block prefix + warp scan), spawn/yaw
//                         reset, manager resets + episode statistics, bounded target rejection sampling on the
//                         valid-location mask, heightmap lookup, command metrics / update, observation head.
//
// Reference lines restated (relative to the reference root):
//   rover_envs/mdp/actions/ackermann_actions.py:226-322        rover_envs/envs/navigation/mdp/rewards.py:14-137
//   rover_envs/envs/navigation/mdp/terminations.py:14-64        .../mdp/observations.py:15-32
//   .../mdp/randomizations.py:12-39                             .../utils/terrains/terrain_importer.py:74-175
//   .../utils/terrains/terrain_utils.py:62-84, 202-223          .../entrypoints/rover_env.py:61-102 (ordering)
// plus the ORBIT manager combine rules and math helpers of SURVEY.md Appendix A.1/A.2.
// fp32 arithmetic follows the reference's operation order; contraction is disabled (__f*_rn) wherever a
// value is compared against a threshold or truncated to an index.
#include <cstring>

#include "common.cuh"
#include "rng.cuh"

namespace rover {

__device__ __forceinline__ float sgnf(float x) { return (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f); }

// torch.remainder(a, 2pi) followed by the (a > pi) fold  (ORBIT wrap_to_pi, A.1)
__device__ __forceinline__ float wrap_to_pi(float a) {
    const float two_pi = 6.2831855f, pi = 3.1415927f;
    float m = fmodf(a, two_pi);
    if (m != 0.f && m < 0.f) m = __fadd_rn(m, two_pi);
    if (m > pi) m = __fsub_rn(m, two_pi);
    return m;
}

struct YawQuat {
    float cw, sz;
};

// ORBIT yaw_quat (A.1)
__device__ __forceinline__ YawQuat yaw_quat(float w, float x, float y, float z) {
    const float siny = __fmul_rn(2.f, __fadd_rn(__fmul_rn(w, z), __fmul_rn(x, y)));
    const float cosy = __fsub_rn(1.f, __fmul_rn(2.f, __fadd_rn(__fmul_rn(y, y), __fmul_rn(z, z))));
    const float half = __fdiv_rn(atan2f(siny, cosy), 2.f);
    const float s = sinf(half), c = cosf(half);
    const float n = fmaxf(sqrtf(__fadd_rn(__fmul_rn(c, c), __fmul_rn(s, s))), 1e-9f);
    return {__fdiv_rn(c, n), __fdiv_rn(s, n)};
}

// ORBIT ArticulationData.heading_w (A.1): atan2 of the rotated x axis
__device__ __forceinline__ float heading_w(float w, float x, float y, float z) {
    const float ty = __fmul_rn(z, 2.f), tz = __fmul_rn(-y, 2.f);  // t = 2 * (xyz x (1,0,0)) = (0, 2z, -2y)
    const float cx = __fsub_rn(__fmul_rn(y, tz), __fmul_rn(z, ty));
    const float cy = __fsub_rn(0.f, __fmul_rn(x, tz));  // z*t.x - x*t.z with t.x = 0
    const float fx = __fadd_rn(1.f, cx);
    const float fy = __fadd_rn(__fmul_rn(w, ty), cy);
    return atan2f(fy, fx);
}

__device__ __forceinline__ float norm2(float x, float y) {
    return sqrtf(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)));
}

// contact_sensor force_matrix_w [B,1,3] of one env: L2 norm over bodies per axis, summed over axes, > 1
template <int kBodies>
__device__ __forceinline__ bool collision_active_fixed(const float* __restrict__ f) {
    // all loads are issued before the first use (one memory round trip); accumulation order = body order
    float v[3 * kBodies];
#pragma unroll
    for (int k = 0; k < 3 * kBodies; ++k) v[k] = __ldg(f + k);
    float sx = 0.f, sy = 0.f, sz = 0.f;
#pragma unroll
    for (int b = 0; b < kBodies; ++b) {
        sx = __fadd_rn(sx, __fmul_rn(v[3 * b], v[3 * b]));
        sy = __fadd_rn(sy, __fmul_rn(v[3 * b + 1], v[3 * b + 1]));
        sz = __fadd_rn(sz, __fmul_rn(v[3 * b + 2], v[3 * b + 2]));
    }
    return __fadd_rn(__fadd_rn(sqrtf(sx), sqrtf(sy)), sqrtf(sz)) > 1.f;
}

__device__ __forceinline__ bool collision_active(const float* __restrict__ f, int num_bodies) {
    if (num_bodies == 14) return collision_active_fixed<14>(f);  // AAU rover: 6 Drive + 4 Steer + 3 Boogie + Body
    float sx = 0.f, sy = 0.f, sz = 0.f;
    for (int b = 0; b < num_bodies; ++b) {
        const float x = __ldg(f + 3 * b), y = __ldg(f + 3 * b + 1), z = __ldg(f + 3 * b + 2);
        sx = __fadd_rn(sx, __fmul_rn(x, x));
        sy = __fadd_rn(sy, __fmul_rn(y, y));
        sz = __fadd_rn(sz, __fmul_rn(z, z));
    }
    return __fadd_rn(__fadd_rn(sqrtf(sx), sqrtf(sy)), sqrtf(sz)) > 1.f;
}

// ---- AckermannAction2.ackermann (ackermann_actions.py:238-322): jp [FL,RL,RR,FR], jv [ML,FL,RL,RR,MR,FR]
__device__ __forceinline__ void ackermann_v2(const RoverMdpParams& P, float lin_p, float ang_p, float* jp, float* jv) {
    float dir = sgnf(lin_p);
    const float turn = sgnf(ang_p);
    if (dir == 0.f) dir = 1.f;                                                  // :255
    const float v = fabsf(lin_p), w = fabsf(ang_p);
    const bool moving = (w != 0.f) || (v != 0.f);                               // :262
    float R = moving ? __fdiv_rn(v, w) : INFINITY;                              // :265-266 (x/0 = inf)
    const float r_min = P.min_radius;                                           // :264
    if (R < r_min) R = r_min;                                                   // :267
    const float half_mw = P.middle_wheel_distance / 2.f, half_fr = P.rear_and_front_wheel_distance / 2.f;
    const float r_ml = __fsub_rn(R, half_mw), r_mr = __fadd_rn(R, half_mw);     // :271-272
    const float r_l = __fsub_rn(R, half_fr), r_r = __fadd_rn(R, half_fr);       // :273-276
    const bool point = R < P.middle_wheel_distance;                             // :278
    const float spin = __fmul_rn(__fadd_rn(v, 1.f), turn);
    const float v_l = point ? -spin : __fmul_rn((w == 0.f) ? v : __fmul_rn(r_l, w), dir);
    const float v_r = point ? spin : __fmul_rn((w == 0.f) ? v : __fmul_rn(r_r, w), dir);
    const float v_ml = point ? -spin : __fmul_rn((w == 0.f) ? v : __fmul_rn(r_ml, w), dir);
    const float v_mr = point ? spin : __fmul_rn((w == 0.f) ? v : __fmul_rn(r_mr, w), dir);
    const float ack = __fmul_rn(atan2f(P.wheelbase_length, r_l), turn);         // :305 (FL radius for all four)
    const float q = 0.78539816339744830962f;
    jv[0] = __fdiv_rn(v_ml, P.wheel_radius);                                    // [ML,FL,RL,RR,MR,FR] :316
    jv[1] = __fdiv_rn(v_l, P.wheel_radius);
    jv[2] = __fdiv_rn(v_l, P.wheel_radius);
    jv[3] = __fdiv_rn(v_r, P.wheel_radius);
    jv[4] = __fdiv_rn(v_mr, P.wheel_radius);
    jv[5] = __fdiv_rn(v_r, P.wheel_radius);
    jp[0] = point ? -q : ack;                                                   // [FL,RL,RR,FR] :317
    jp[1] = point ? q : ack;
    jp[2] = point ? -q : ack;
    jp[3] = point ? q : ack;
}

// ---- AckermannAction.ackermann (ackermann_actions.py:91-158): jp [FL,FR,RL,RR], jv [FL,FR,ML,MR,RL,RR]
__device__ __forceinline__ void ackermann_v1(float lin, float ang, float* jp, float* jv) {
    const float wx[6] = {-0.385f, 0.385f, -0.447f, 0.447f, -0.385f, 0.385f};    // :97-108 (x right, y forward)
    const float wy[6] = {0.438f, 0.438f, 0.f, 0.f, -0.411f, -0.411f};
    float p = copysignf(__fdiv_rn(lin, ang), -ang);                             // :117-118
    p = (fabsf(p) > 0.45f) ? p : 0.f;                                           // :122
    const float lin2 = (p != 0.f) ? lin : 0.f;                                  // :123
    const float w_lin = copysignf(ang, lin2);                                   // :134
    float steer[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const float dx = __fsub_rn(p, wx[k]), dy = __fsub_rn(0.f, wy[k]);
        const float dist = sqrtf(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));   // :127
        const float side = (k & 1) ? 1.f : -1.f;                                // :130-131
        const float av = (lin2 != 0.f) ? w_lin : __fmul_rn(ang, side);          // :136-137
        float vel = __fmul_rn(dist, av);                                        // :141
        vel = (dist > 1000.f) ? lin2 : vel;                                     // :144
        jv[k] = __fdiv_rn(vel, 0.2f);                                           // :147
        float s = atan2f(wy[k], __fsub_rn(wx[k], p));                           // :149-154
        if (s < -1.57f) s = __fadd_rn(s, 3.1415927f);                           // :155
        if (s > 1.57f) s = __fsub_rn(s, 3.1415927f);                            // :156
        steer[k] = s;
    }
    jp[0] = steer[0];                                                           // :158
    jp[1] = steer[1];
    jp[2] = steer[4];
    jp[3] = steer[5];
}

// ---- ackermann() used by AckermannAction3 (ackermann_actions.py:423-505): jp [FL,FR,RL,RR], jv [FL,FR,ML,MR,RL,RR]
__device__ __forceinline__ void ackermann_v3(const RoverMdpParams& P, float lin_p, float ang_p, float* jp, float* jv) {
    float dir = sgnf(lin_p);
    const float turn = sgnf(ang_p);
    if (dir == 0.f) dir = 1.f;
    const float v = fabsf(lin_p), w = fabsf(ang_p);
    const bool moving = (w != 0.f) || (v != 0.f);
    const float R = moving ? __fdiv_rn(v, w) : INFINITY;                        // :443-444 (no clamp, :445)
    const float hm = __fmul_rn(P.middle_wheel_distance / 2.f, turn), hf = __fmul_rn(P.rear_and_front_wheel_distance / 2.f, turn);
    const float r_ml = __fsub_rn(R, hm), r_mr = __fadd_rn(R, hm);               // :449-450
    const float r_l = __fsub_rn(R, hf), r_r = __fadd_rn(R, hf);                 // :451-454
    const bool point = R < P.min_radius;                                        // :461
    const float spin = __fmul_rn(__fadd_rn(v, 1.f), turn);
    const float v_l = point ? -spin : __fmul_rn((w == 0.f) ? v : __fmul_rn(r_l, w), dir);
    const float v_r = point ? spin : __fmul_rn((w == 0.f) ? v : __fmul_rn(r_r, w), dir);
    const float v_ml = point ? -spin : __fmul_rn((w == 0.f) ? v : __fmul_rn(r_ml, w), dir);
    const float v_mr = point ? spin : __fmul_rn((w == 0.f) ? v : __fmul_rn(r_mr, w), dir);
    const float half_wl = P.wheelbase_length / 2.f;
    const float y_front = __fsub_rn(half_wl, P.offset_lin), y_rear = __fadd_rn(half_wl, P.offset_lin);  // :488-497
    const float q = 0.78539816339744830962f;
    jp[0] = point ? -q : __fmul_rn(atan2f(y_front, r_l), turn);                 // FL
    jp[1] = point ? q : __fmul_rn(atan2f(y_front, r_r), turn);                  // FR
    jp[2] = point ? q : __fmul_rn(atan2f(y_rear, r_l), -turn);                  // RL
    jp[3] = point ? -q : __fmul_rn(atan2f(y_rear, r_r), -turn);                 // RR
    const float diam = P.wheel_diameter;                                        // :503 (float)(wheel_radius * 2)
    jv[0] = __fdiv_rn(v_l, diam);
    jv[1] = __fdiv_rn(v_r, diam);
    jv[2] = __fdiv_rn(v_ml, diam);
    jv[3] = __fdiv_rn(v_mr, diam);
    jv[4] = __fdiv_rn(v_l, diam);
    jv[5] = __fdiv_rn(v_r, diam);
}

__device__ __forceinline__ void ackermann_dispatch(const RoverMdpParams& P, float lin_p, float ang_p, float* jp, float* jv) {
    if (P.action_variant == 1) ackermann_v1(lin_p, ang_p, jp, jv);
    else if (P.action_variant == 3) ackermann_v3(P, lin_p, ang_p, jp, jv);
    else ackermann_v2(P, lin_p, ang_p, jp, jv);
}

__global__ void ackermann_kernel(const float* __restrict__ actions, int n, const __grid_constant__ RoverMdpParams P,
                                 float* __restrict__ processed, float* __restrict__ jp, float* __restrict__ jv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float lin_p = __fadd_rn(__fmul_rn(actions[2 * i], P.scale_lin), P.offset_lin);
    const float ang_p = __fadd_rn(__fmul_rn(actions[2 * i + 1], P.scale_ang), P.offset_ang);
    if (processed) {
        processed[2 * i] = lin_p;
        processed[2 * i + 1] = ang_p;
    }
    ackermann_dispatch(P, lin_p, ang_p, jp + 4 * (size_t)i, jv + 6 * (size_t)i);
}

// The per-env work of the pre-step (one thread per env); returns the env's reset flag.
__device__ __forceinline__ bool pre_step_env(int i, const float* __restrict__ new_actions, const float* __restrict__ force,
                                             int n, const RoverMdpParams& P, const RoverMdpState& S, const RoverMdpOut& O,
                                             int phases) {
    bool reset = false;
    if (i < n) {
        float2 a_old, a;
        if (phases & ROVER_PRE_ACTIONS) {
        // ---- ActionManager.process_action: prev <- action <- new; term.process_actions (ackermann_actions.py:226-229)
        a_old = reinterpret_cast<const float2*>(S.action)[i];
        a = reinterpret_cast<const float2*>(new_actions)[i];
        reinterpret_cast<float2*>(S.prev_action)[i] = a_old;
        reinterpret_cast<float2*>(S.action)[i] = a;
        const float lin_p = __fadd_rn(__fmul_rn(a.x, P.scale_lin), P.offset_lin);
        const float ang_p = __fadd_rn(__fmul_rn(a.y, P.scale_ang), P.offset_ang);
        reinterpret_cast<float2*>(O.processed_actions)[i] = make_float2(lin_p, ang_p);

        ackermann_dispatch(P, lin_p, ang_p, O.joint_pos + 4 * (size_t)i, O.joint_vel + 6 * (size_t)i);
        } else {
            a = reinterpret_cast<const float2*>(S.action)[i];
            a_old = reinterpret_cast<const float2*>(S.prev_action)[i];
        }

        if (phases & ROVER_PRE_TERMS) {
        // ---- counters (rover_env.py:79)
        const long long ep = S.episode_length_buf[i] + 1;
        S.episode_length_buf[i] = ep;

        // ---- shared quantities of the PREVIOUS command (rover_env.py:82-86 run before the command update)
        const float bx = S.pos_cmd_b[3 * (size_t)i], by = S.pos_cmd_b[3 * (size_t)i + 1];
        const float d = norm2(bx, by);
        const float ang = atan2f(by, bx);
        const bool coll = collision_active(force + (size_t)i * P.num_bodies * 3, P.num_bodies);
        const float max_len = (float)P.max_episode_length;

        // ---- terminations (terminations.py:14-64, ORBIT mdp.time_out)
        const bool t_out = ep >= (long long)P.max_episode_length;
        const bool t_succ = d < P.reached_threshold;
        const bool t_far = d > P.far_threshold;
        const bool terminated = t_succ || t_far || coll;
        reset = t_out || terminated;
        O.terminated[i] = terminated;
        O.truncated[i] = t_out;
        reinterpret_cast<uchar4*>(O.term_flags)[i] = make_uchar4(t_out, t_succ, t_far, coll);
        O.reset_flags[i] = reset;

        // ---- rewards (rewards.py:14-137), RewardManager: value * weight * dt, summed in declaration order
        float val[ROVER_NUM_REWARD_TERMS];
        val[0] = __fdiv_rn(__fdiv_rn(1.f, __fadd_rn(1.f, __fmul_rn(__fmul_rn(0.11f, d), d))), max_len);
        val[1] = t_succ ? __fdiv_rn((float)((long long)P.max_episode_length - ep), max_len) : 0.f;
        {
            const float d_lin = __fmul_rn(__fsub_rn(a.y, a_old.y), 3.f);
            const float d_ang = __fmul_rn(__fsub_rn(a.x, a_old.x), 3.f);
            float p_ang = (d_ang > 0.05f) ? __fmul_rn(d_ang, d_ang) : 0.f;
            float p_lin = (d_lin > 0.05f) ? __fmul_rn(d_lin, d_lin) : 0.f;
            p_ang = __fmul_rn(p_ang, p_ang);
            p_lin = __fmul_rn(p_lin, p_lin);
            val[2] = __fdiv_rn(__fadd_rn(p_ang, p_lin), max_len);
        }
        val[3] = (fabsf(ang) > 2.f) ? __fdiv_rn(fabsf(ang), max_len) : 0.f;
        val[4] = (a.x < 0.f) ? (float)(1.0 / (double)P.max_episode_length) : 0.f;
        val[5] = coll ? 1.f : 0.f;
        val[6] = t_far ? 1.f : 0.f;
        float total = 0.f;
        float* sums = S.episode_sums + ROVER_NUM_REWARD_TERMS * (size_t)i;
        float* tr = O.term_rewards + ROVER_NUM_REWARD_TERMS * (size_t)i;
        float* tv = O.term_values + ROVER_NUM_REWARD_TERMS * (size_t)i;
#pragma unroll
        for (int k = 0; k < ROVER_NUM_REWARD_TERMS; ++k) {
            const float c = __fmul_rn(__fmul_rn(val[k], P.weight[k]), P.step_dt);
            total = __fadd_rn(total, c);
            sums[k] = __fadd_rn(sums[k], c);
            tr[k] = c;
            tv[k] = val[k];
        }
        O.reward[i] = total;
        }
    }
    return reset;
}

__global__ void __launch_bounds__(ROVER_MDP_BLOCK)
mdp_pre_step_kernel(const float* __restrict__ new_actions, const float* __restrict__ force, int n,
                    const __grid_constant__ RoverMdpParams P, const __grid_constant__ RoverMdpState S,
                    const __grid_constant__ RoverMdpOut O, int phases) {
    const int i = blockIdx.x * ROVER_MDP_BLOCK + threadIdx.x;
    const bool reset = pre_step_env(i, new_actions, force, n, P, S, O, phases);
    if (phases & ROVER_PRE_TERMS) {
        const int cnt = __syncthreads_count(reset);
        if (threadIdx.x == 0) O.block_reset_counts[blockIdx.x] = cnt;
    }
}

// --------------------------------------------------------------------------------------------------------------

struct Tables {
    const float* __restrict__ heightmap;
    const uint8_t* __restrict__ safe_mask;
    int H, W;
    float offx, offy, res;
    const float* __restrict__ spawn;
    int n_spawns;
};

// terrain_utils.py:75-81 / :211-218: cell = trunc(xy / res + (min_x, min_y)), clamped (offset ADDED, sic)
__device__ __forceinline__ void terrain_cell(const Tables& T, float x, float y, int& col, int& row) {
    const float sx = __fadd_rn(__fdiv_rn(x, T.res), T.offx);
    const float sy = __fadd_rn(__fdiv_rn(y, T.res), T.offy);
    const long long cx = (long long)fminf(fmaxf(sx, -1.0e18f), 1.0e18f);  // .long(): truncation toward zero
    const long long cy = (long long)fminf(fmaxf(sy, -1.0e18f), 1.0e18f);
    col = (int)min(max(cx, 0LL), (long long)(T.W - 1));
    row = (int)min(max(cy, 0LL), (long long)(T.H - 1));
}

// CommandTerm._resample + _resample_command + sample_new_targets (terrain_importer.py:74-95, 134-175).
// The reference's rejection loop is sequential (one host sync per round); here the candidates of a batch of 8 rounds
// are generated together and their mask bytes fetched concurrently (one memory round trip per batch instead of per
// round), then the first valid round wins -- the same candidate the sequential loop would have accepted.  The height
// under every candidate is fetched in the same round trip (speculatively: 8 loads for the ~5 % of envs that resample),
// which takes the heightmap lookup off the dependent chain rank -> spawn row -> mask -> height.
// The random variates of the reset path: explicit arrays (parity tests: oracle and kernel read the same numbers) or the
// counter-based generator of rng.cuh evaluated in registers (rng != nullptr: {seed, step} in device memory, the step
// counter advanced by the launch itself, so the launch can sit in a CUDA graph).
struct VariatesDev {
    const long long* __restrict__ spawn_perm;
    const float* __restrict__ yaw_u;
    const float* __restrict__ heading_u;
    const float* __restrict__ theta_u;
    unsigned long long* rng;
    int n_rounds;
};

template <bool kRng>
__device__ __forceinline__ bool resample_command(int i, const RoverMdpParams& P, const RoverMdpState& S, const Tables& T,
                                                 float ox, float oy, const float* __restrict__ theta_u,
                                                 const float (&theta0)[8], const RngKey& key, int n_rounds,
                                                 float heading_u, float& cx, float& cy, float& cz, float& chead) {
    const float pi_f = 3.1415927f;  // torch.pi as fp32; the reference computes rand * 2 * pi left to right
    constexpr int kBatch = 8;
    float x = 0.f, y = 0.f, z = 0.f;
    bool bad = true;
    for (int r0 = 0; r0 < n_rounds && bad; r0 += kBatch) {
        float u[kBatch], xs[kBatch], ys[kBatch];
        int cols[kBatch], rows[kBatch];
        uint8_t m[kBatch];
        float hz[kBatch];
        if (kRng) {  // rounds 4q .. 4q+3 = Philox stream 1 + q of this env
#pragma unroll
            for (int q = 0; q < kBatch / 4; ++q) {
                uint32_t w[4];
                rng_env_stream(key, (uint32_t)i, 1u + (uint32_t)(r0 / 4 + q), w);
#pragma unroll
                for (int k = 0; k < 4; ++k) u[4 * q + k] = u01(w[k]);
            }
        } else {
#pragma unroll
            for (int k = 0; k < kBatch; ++k)  // the first batch was prefetched by the caller
                u[k] = (r0 == 0) ? theta0[k] : ((r0 + k < n_rounds) ? __ldg(theta_u + (size_t)i * n_rounds + r0 + k) : 0.f);
        }
#pragma unroll
        for (int k = 0; k < kBatch; ++k) {
            const float th = __fmul_rn(__fmul_rn(u[k], 2.f), pi_f);                      // :169
            xs[k] = __fadd_rn(__fmul_rn(cosf(th), P.target_distance), ox);               // :172
            ys[k] = __fadd_rn(__fmul_rn(sinf(th), P.target_distance), oy);               // :173
            terrain_cell(T, xs[k], ys[k], cols[k], rows[k]);
            m[k] = __ldg(T.safe_mask + (size_t)rows[k] * T.W + cols[k]);                 // :220
            hz[k] = __ldg(T.heightmap + (size_t)rows[k] * T.W + cols[k]);                // :154, used if round k wins
        }
#pragma unroll
        for (int k = 0; k < kBatch; ++k) {
            if (bad && r0 + k < n_rounds) {  // sequential semantics: the first valid round, else the last tried
                x = xs[k], y = ys[k], z = hz[k];
                bad = m[k] == 1;
            }
        }
    }
    cx = x;
    cy = y;
    cz = z;                                                                  // :154 (+ default_root_state z = 0)
    chead = __fadd_rn(__fmul_rn(heading_u, __fsub_rn(P.heading_hi, P.heading_lo)), P.heading_lo);  // uniform_(lo, hi)
    S.time_left[i] = P.resampling_time;
    return bad;
}

constexpr int kStats = ROVER_STATS_LEN;

struct StatsExchangeDev {
    void* const* peer_mailbox;
    double* cumulative;
    unsigned long long* sequence;
    int rank, world;  // world == 0: no exchange
};

// look-back descriptor of the fused step: [epoch : 30 | status : 2 | value : 32]
constexpr unsigned long long kDescAggregate = 1ull, kDescPrefix = 2ull;
__device__ __forceinline__ unsigned long long make_lookback(unsigned epoch, unsigned long long status, unsigned value) {
    return ((unsigned long long)(epoch & 0x3fffffffu) << 34) | (status << 32) | value;
}

// The block-level work of the post-step.  kFused = false: the reset flags and the per-block reset counts come from the
// pre-step launch.  kFused = true (rover_mdp_step): `reset_in` comes from pre_step_env of the same thread and the rank of
// the block's first reset env from a decoupled look-back over the blocks' reset counts (no second launch).
template <bool kFused, bool kRng>
__device__ __forceinline__ void post_step_block(int bid, int n_blocks, bool reset_in, float* __restrict__ root_pos_w,
                                                float* __restrict__ root_quat_w, int n, const RoverMdpParams& P,
                                                const RoverMdpState& S, const RoverMdpOut& O, const Tables& T,
                                                const VariatesDev& V, long long* __restrict__ out_spawn_index,
                                                float* __restrict__ block_stats, unsigned int* __restrict__ done_counter,
                                                float* __restrict__ stats, float* __restrict__ log_out,
                                                float* __restrict__ obs, int obs_stride, int phases,
                                                const StatsExchangeDev& X, unsigned long long* __restrict__ lookback,
                                                unsigned epoch) {
    const long long* __restrict__ spawn_perm = V.spawn_perm;
    const float* __restrict__ yaw_u = V.yaw_u;
    const float* __restrict__ heading_u = V.heading_u;
    const float* __restrict__ theta_u = V.theta_u;
    const int n_rounds = V.n_rounds;
    // {seed, step}: one uniform load per thread (L2 / L1 hit after the first); the step word is advanced by the last
    // block of the launch, after every block has read it
    RngKey key = make_rng_key(0ull, 0ull);
    if (kRng) key = make_rng_key(V.rng[0], *reinterpret_cast<volatile unsigned long long*>(V.rng + 1));
    __shared__ int warp_cnt[ROVER_MDP_BLOCK / 32];
    __shared__ int block_base;
    constexpr int kRedRows = ROVER_MDP_BLOCK / 4;  // 16 row groups of the last-block reduction (>= warps per block)
    __shared__ float red[kRedRows][kStats];
    __shared__ bool is_last;
    const int i = bid * ROVER_MDP_BLOCK + threadIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const bool valid = i < n;
    const bool reset = kFused ? (valid && reset_in) : (valid && O.reset_flags[i] != 0);

    // ---- every load that does not depend on the reset rank is issued up front, so that the kernel pays one memory
    //      round trip for them instead of one per dependent stage (this launch is latency-bound, not bandwidth-bound)
    float px = 0.f, py = 0.f, pz = 0.f, cwx = 0.f, cwy = 0.f, cwz = 0.f, chead = 0.f, time_left = 0.f;
    float yaw_var = 0.f, heading_var = 0.f, theta0[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float4 q = make_float4(1.f, 0.f, 0.f, 0.f);
    float2 act = make_float2(0.f, 0.f);
    if (valid) {
        px = root_pos_w[3 * (size_t)i], py = root_pos_w[3 * (size_t)i + 1], pz = root_pos_w[3 * (size_t)i + 2];
        q = reinterpret_cast<float4*>(root_quat_w)[i];  // (w,x,y,z)
        cwx = S.pos_cmd_w[3 * (size_t)i], cwy = S.pos_cmd_w[3 * (size_t)i + 1], cwz = S.pos_cmd_w[3 * (size_t)i + 2];
        chead = S.heading_cmd_w[i];
        act = reinterpret_cast<float2*>(S.action)[i];
        time_left = S.time_left[i];
        if (!kRng) {
            yaw_var = __ldg(yaw_u + i);
            heading_var = __ldg(heading_u + i);
#pragma unroll
            for (int k = 0; k < 8; ++k) theta0[k] = (k < n_rounds) ? __ldg(theta_u + (size_t)i * n_rounds + k) : 0.f;
        }
    }

    // ---- explicit variates: rank of this env among the reset envs in ascending env order (== reset_buf.nonzero() order),
    //      the index into spawn_perm (randperm(len)[:K] assigned in that order, randomizations.py:22).  The in-kernel
    //      generator needs no rank: its spawn draw is a keyed permutation of [0, n_spawns) evaluated at the ENV ID -- K
    //      distinct envs get K distinct rows, and a uniformly random permutation evaluated at K distinct points has
    //      exactly the distribution of randperm(len)[:K] -- so the cross-block prefix (a dependent load in front of the
    //      reset chain; a look-back in the single-launch step) disappears.
    int rank = 0;
    if constexpr (!kRng) {
        const unsigned ballot = __ballot_sync(0xffffffffu, reset);
        if (lane == 0) warp_cnt[wid] = __popc(ballot);
        if (!kFused) {
            if (wid == 0) {
                int acc = 0;
                for (int b = lane; b < bid; b += 32) acc += O.block_reset_counts[b];
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                if (lane == 0) block_base = acc;
            }
            __syncthreads();
        } else {
            __syncthreads();
            if (wid == 0) {
                // decoupled look-back (single pass): publish this block's count, walk back over the predecessors' descriptors
                // until one carries an inclusive prefix, publish our own.  Logical block ids are tickets, so every
                // predecessor has started; descriptors are tagged with the launch epoch, so nothing is re-zeroed.
                int cnt = 0;
                for (int w = 0; w < ROVER_MDP_BLOCK / 32; ++w) cnt += warp_cnt[w];
                volatile unsigned long long* desc = lookback;
                int base = 0;
                if (bid > 0) {
                    if (lane == 0) desc[bid] = make_lookback(epoch, kDescAggregate, (unsigned)cnt);
                    int look = bid - 1;
                    while (true) {
                        const int idx = look - lane;
                        unsigned long long d = make_lookback(epoch, kDescPrefix, 0u);  // before block 0: prefix 0
                        if (idx >= 0) {
                            do {
                                d = desc[idx];
                            } while ((unsigned)(d >> 34) != (epoch & 0x3fffffffu) || ((d >> 32) & 3ull) == 0ull);
                        }
                        const bool is_prefix = ((d >> 32) & 3ull) == kDescPrefix;
                        const unsigned pmask = __ballot_sync(0xffffffffu, is_prefix);
                        const int first = pmask ? __ffs(pmask) - 1 : 32;  // closest predecessor with an inclusive prefix
                        int v = (lane <= first) ? (int)(unsigned)(d & 0xffffffffull) : 0;
                        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                        base += v;
                        if (pmask) break;
                        look -= 32;
                    }
                }
                if (lane == 0) {
                    __threadfence();
                    desc[bid] = make_lookback(epoch, kDescPrefix, (unsigned)(base + cnt));
                    block_base = base;
                }
            }
            __syncthreads();
        }
        rank = block_base + __popc(ballot & ((1u << lane) - 1u));
        for (int w = 0; w < wid; ++w) rank += warp_cnt[w];

    }

    float st[kStats];
#pragma unroll
    for (int k = 0; k < kStats; ++k) st[k] = 0.f;

    if (valid) {
        long long spawn_idx = -1;
        bool cmd_dirty = false;
        bool origin_known = false;  // env origin of a freshly spawned env stays in registers (no store -> load round trip)
        float org_x = 0.f, org_y = 0.f;

        if (kRng && (reset || time_left <= P.step_dt)) {  // the only envs that consume variates this step
            uint32_t w[4];
            rng_env_stream(key, (uint32_t)i, 0u, w);
            yaw_var = u01(w[0]);
            heading_var = u01(w[1]);
        }
        if (reset && (phases & ROVER_PHASE_SPAWN)) {
            // -- reset_root_state_rover (randomizations.py:12-39)
            if (kRng) spawn_idx = (long long)spawn_perm_at(make_spawn_perm_key(key, (uint32_t)T.n_spawns), (uint32_t)i);
            else spawn_idx = __ldg(spawn_perm + rank);
            const float* sp = T.spawn + 3 * (size_t)spawn_idx;
            px = __ldg(sp);
            py = __ldg(sp + 1);
            pz = __fadd_rn(__ldg(sp + 2), P.spawn_z_offset);
            const float angle = __fmul_rn(__fmul_rn(yaw_var, 2.f), 3.1415927f);
            const float half = __fdiv_rn(angle, 2.f);
            q = make_float4(cosf(half), 0.f, 0.f, sinf(half));
            org_x = px, org_y = py, origin_known = true;
            S.env_origins[3 * (size_t)i] = px;
            S.env_origins[3 * (size_t)i + 1] = py;
            S.env_origins[3 * (size_t)i + 2] = pz;
            root_pos_w[3 * (size_t)i] = px;
            root_pos_w[3 * (size_t)i + 1] = py;
            root_pos_w[3 * (size_t)i + 2] = pz;
            reinterpret_cast<float4*>(root_quat_w)[i] = q;
        }
        if (reset && (phases & ROVER_PHASE_MANAGERS)) {
            // -- ActionManager.reset
            act = make_float2(0.f, 0.f);
            reinterpret_cast<float2*>(S.action)[i] = act;
            reinterpret_cast<float2*>(S.prev_action)[i] = act;
            // -- RewardManager.reset: episodic sums of reset envs -> stats, then zero
            float* sums = S.episode_sums + ROVER_NUM_REWARD_TERMS * (size_t)i;
#pragma unroll
            for (int k = 0; k < ROVER_NUM_REWARD_TERMS; ++k) {
                st[k] = sums[k];
                sums[k] = 0.f;
            }
            // -- TerminationManager.reset: per-term counts
            const uchar4 tf = reinterpret_cast<const uchar4*>(O.term_flags)[i];
            st[7] = tf.x;
            st[8] = tf.y;
            st[9] = tf.z;
            st[10] = tf.w;
            // -- CommandTerm.reset: metrics -> stats, zero, counter = 0, then _resample (counter += 1)
            st[11] = S.err_pos[i];
            st[12] = S.err_heading[i];
            st[13] = 1.f;
            S.err_pos[i] = 0.f;
            S.err_heading[i] = 0.f;
            S.command_counter[i] = 0;
            S.episode_length_buf[i] = 0;
        }
        if (reset && (phases & ROVER_PHASE_RESAMPLE)) {
            // -- CommandTerm._resample: time_left, counter += 1, _resample_command around the (new) env origin
            const float ox = origin_known ? org_x : S.env_origins[3 * (size_t)i];
            const float oy = origin_known ? org_y : S.env_origins[3 * (size_t)i + 1];
            const bool exhausted = resample_command<kRng>(i, P, S, T, ox, oy, theta_u, theta0, key, n_rounds, heading_var,
                                                          cwx, cwy, cwz, chead);
            st[14] = exhausted ? 1.f : 0.f;
            S.command_counter[i] += 1;
            time_left = P.resampling_time;
            cmd_dirty = true;
        }
        if (out_spawn_index) out_spawn_index[i] = spawn_idx;

        // -- CommandManager.compute(dt): metrics, time_left, time-based resample, _update_command
        const float hw = heading_w(q.x, q.y, q.z, q.w);
        if (phases & ROVER_PHASE_METRICS) {
            const float ex = __fsub_rn(cwx, px), ey = __fsub_rn(cwy, py), ez = __fsub_rn(cwz, pz);
            S.err_pos[i] = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez)));
            S.err_heading[i] = fabsf(wrap_to_pi(__fsub_rn(chead, hw)));
        }
        if (phases & ROVER_PHASE_TIME) time_left = __fsub_rn(time_left, P.step_dt);
        if ((phases & ROVER_PHASE_TIME) && time_left <= 0.f) {
            const float ox = origin_known ? org_x : S.env_origins[3 * (size_t)i];
            const float oy = origin_known ? org_y : S.env_origins[3 * (size_t)i + 1];
            // (a reset resample above and this one never meet in one step: the reset sets time_left = resampling_time,
            // which the launcher requires to exceed step_dt -- so sharing the env's variates between them is safe)
            const bool exhausted = resample_command<kRng>(i, P, S, T, ox, oy, theta_u, theta0, key, n_rounds, heading_var,
                                                          cwx, cwy, cwz, chead);
            st[14] += exhausted ? 1.f : 0.f;
            st[15] = 1.f;
            S.command_counter[i] += 1;
            time_left = P.resampling_time;
            cmd_dirty = true;
        }
        if (phases & (ROVER_PHASE_TIME | ROVER_PHASE_RESAMPLE)) S.time_left[i] = time_left;
        if (cmd_dirty) {
            S.pos_cmd_w[3 * (size_t)i] = cwx;
            S.pos_cmd_w[3 * (size_t)i + 1] = cwy;
            S.pos_cmd_w[3 * (size_t)i + 2] = cwz;
            S.heading_cmd_w[i] = chead;
        }
        float pbx = S.pos_cmd_b[3 * (size_t)i], pby = S.pos_cmd_b[3 * (size_t)i + 1];
        if (phases & ROVER_PHASE_COMMAND) {
        // _update_command (terrain_importer.py:97-101): quat_rotate_inverse(yaw_quat(q), target - root)
        const float vx = __fsub_rn(cwx, px), vy = __fsub_rn(cwy, py), vz = __fsub_rn(cwz, pz);
        const YawQuat yq = yaw_quat(q.x, q.y, q.z, q.w);
        const float k = __fsub_rn(__fmul_rn(2.f, __fmul_rn(yq.cw, yq.cw)), 1.f);
        const float b_x = __fmul_rn(__fmul_rn(-__fmul_rn(yq.sz, vy), yq.cw), 2.f);
        const float b_y = __fmul_rn(__fmul_rn(__fmul_rn(yq.sz, vx), yq.cw), 2.f);
        const float dot = __fmul_rn(yq.sz, vz);
        const float c_z = __fmul_rn(__fmul_rn(yq.sz, dot), 2.f);
        pbx = __fsub_rn(__fmul_rn(vx, k), b_x);
        pby = __fsub_rn(__fmul_rn(vy, k), b_y);
        const float pbz = __fadd_rn(__fmul_rn(vz, k), c_z);
        S.pos_cmd_b[3 * (size_t)i] = pbx;
        S.pos_cmd_b[3 * (size_t)i + 1] = pby;
        S.pos_cmd_b[3 * (size_t)i + 2] = pbz;
        S.heading_cmd_b[i] = wrap_to_pi(__fsub_rn(chead, hw));
        }

        // -- observation head (rover_env_cfg.py:103-112): last_action, distance * 0.11, angle / pi
        if (obs && (phases & ROVER_PHASE_OBS)) {
            float* o = obs + (size_t)i * obs_stride;
            o[0] = act.x;
            o[1] = act.y;
            o[2] = __fmul_rn(norm2(pbx, pby), P.obs_distance_scale);
            o[3] = __fmul_rn(atan2f(pby, pbx), P.obs_heading_scale);
        }
    }

    // ---- deterministic episode statistics: warp shuffle -> block -> last block sums the partials in order
#pragma unroll
    for (int k = 0; k < kStats; ++k) {
        float v = st[k];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) red[wid][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < kStats) {
        float v = 0.f;
        for (int w = 0; w < ROVER_MDP_BLOCK / 32; ++w) v += red[w][threadIdx.x];
        block_stats[(size_t)bid * kStats + threadIdx.x] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(done_counter, 1u) == (unsigned)n_blocks - 1u);
    __syncthreads();
    if (is_last) {
        // Fixed summation tree (deterministic, same result for the same inputs whatever the block schedule): thread t
        // owns the statistics quad c = t % 4 of the block rows g, g + 16, ... (g = t / 4): float4 loads, kUnroll of them
        // in flight per thread, so the 16 KB of partials (L2) cost one or two round trips instead of a chain of them;
        // the 16 row groups are then combined in order of g.
        static_assert(ROVER_MDP_BLOCK == 64 && kStats == 16, "last-block reduction layout");
        constexpr int kGroups = ROVER_MDP_BLOCK / 4, kUnroll = 16;
        const int c4 = threadIdx.x & 3, g = threadIdx.x >> 2;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (unsigned b0 = g; b0 < (unsigned)n_blocks; b0 += kGroups * kUnroll) {
            float4 x[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const unsigned bb = b0 + u * kGroups;
                x[u] = bb < (unsigned)n_blocks ? __ldcg(reinterpret_cast<const float4*>(block_stats + (size_t)bb * kStats) + c4)
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) acc.x += x[u].x, acc.y += x[u].y, acc.z += x[u].z, acc.w += x[u].w;
        }
        __syncthreads();  // red[] is free: every thread passed the block-level reduction above
        red[g][4 * c4 + 0] = acc.x, red[g][4 * c4 + 1] = acc.y, red[g][4 * c4 + 2] = acc.z, red[g][4 * c4 + 3] = acc.w;
        __syncthreads();
        constexpr int G = kGroups;
        double total = 0.0;
        if (threadIdx.x < kStats) {
            float t = 0.f;
            for (int q = 0; q < G; ++q) t += red[q][threadIdx.x];
            stats[threadIdx.x] += t;
            // extras["log"] of the ORBIT managers' reset() (A.2), refreshed only by a launch that reset something -- the
            // reference calls _reset_idx only then (rover_env.py:89-91), so the previous values stay visible otherwise:
            //   Episode Reward/<term> = mean(episode_sums[ids]) / max_episode_length_s, Episode Termination/<term> =
            //   count, Metrics/target_pose/<m> = mean over ids; [13] = number of resets, [14], [15] as in stats
            const float cnt = __shfl_sync(0xffffu, t, 13);  // the 16 statistics live in lanes 0..15 of warp 0
            if (log_out != nullptr && (phases & ROVER_PHASE_MANAGERS) && cnt > 0.f) {
                const int k = (int)threadIdx.x;
                float v = t;
                if (k < ROVER_NUM_REWARD_TERMS) v = __fdiv_rn(__fdiv_rn(t, cnt), P.episode_length_s);
                else if (k == 11 || k == 12) v = __fdiv_rn(t, cnt);
                log_out[k] = v;
            }
            if (threadIdx.x == 0) {
                *done_counter = 0u;  // re-arm for the next launch
                if (kRng) V.rng[1] = V.rng[1] + 1ull;  // next launch = next step of the variate streams
                if (kFused && !kRng) {  // fused step, explicit variates: next launch = next epoch, tickets from 0 again
                    lookback[n_blocks] = 0ull;
                    lookback[n_blocks + 1] = (unsigned long long)(epoch + 1u);
                }
            }
            if (X.world > 0) {
                total = X.cumulative[threadIdx.x] + (double)t;
                X.cumulative[threadIdx.x] = total;
            }
        }
        // ---- multi-GPU: publish the running totals into this rank's slot of every rank's mailbox (peer stores over
        //      NVLink).  A slot holds a sequence number and two value buffers; the writer fills buffer (seq + 1) & 1 in
        //      every mailbox, fences once, then stores seq + 1 everywhere -- one system-scope fence per step, all peers
        //      in flight together.  No collective, no extra launch.
        if (X.world > 0) {
            __shared__ double pub[kStats];
            if (threadIdx.x < kStats) pub[threadIdx.x] = total;
            __syncthreads();
            const unsigned long long seq = *X.sequence + 1ull;
            for (int e = threadIdx.x; e < X.world * kStats; e += ROVER_MDP_BLOCK) {
                const int p = e / kStats, k = e % kStats;
                unsigned char* slot = static_cast<unsigned char*>(X.peer_mailbox[p]) + (size_t)X.rank * ROVER_MAILBOX_SLOT_BYTES;
                reinterpret_cast<volatile double*>(slot + 8)[(seq & 1ull) * kStats + k] = pub[k];
            }
            __threadfence_system();
            __syncthreads();
            for (int p = threadIdx.x; p < X.world; p += ROVER_MDP_BLOCK) {
                unsigned char* slot = static_cast<unsigned char*>(X.peer_mailbox[p]) + (size_t)X.rank * ROVER_MAILBOX_SLOT_BYTES;
                *reinterpret_cast<volatile unsigned long long*>(slot) = seq;
            }
            if (threadIdx.x == 0) *X.sequence = seq;
        }
    }
}

template <bool kRng>
__global__ void __launch_bounds__(ROVER_MDP_BLOCK)
mdp_post_step_kernel(float* __restrict__ root_pos_w, float* __restrict__ root_quat_w, int n,
                     const __grid_constant__ RoverMdpParams P, const __grid_constant__ RoverMdpState S,
                     const __grid_constant__ RoverMdpOut O, const __grid_constant__ Tables T,
                     const __grid_constant__ VariatesDev V, long long* __restrict__ out_spawn_index,
                     float* __restrict__ block_stats, unsigned int* __restrict__ done_counter, float* __restrict__ stats,
                     float* __restrict__ log_out, float* __restrict__ obs, int obs_stride, int phases,
                     const __grid_constant__ StatsExchangeDev X) {
    post_step_block<false, kRng>((int)blockIdx.x, (int)gridDim.x, false, root_pos_w, root_quat_w, n, P, S, O, T, V,
                                 out_spawn_index, block_stats, done_counter, stats, log_out, obs, obs_stride, phases, X,
                                 nullptr, 0u);
}

// rover_mdp_step: pre-step + post-step of one env block in ONE launch (the reset rank comes from a look-back instead of a
// second launch; the pre-step's outputs of an env are consumed by the same thread).  lookback: [n_blocks] descriptors,
// then the ticket counter and the epoch (both maintained by the kernel itself, so the launch can sit in a CUDA graph).
// Logical block ids are ALWAYS tickets: the look-back spins on predecessors, which is only safe if every predecessor
// has started -- blockIdx order guarantees that on an idle GPU only, a ticket taken at block start guarantees it under
// MPS, concurrent kernels or a partitioned device as well (one atomic per block).
template <bool kRng>
__global__ void __launch_bounds__(ROVER_MDP_BLOCK)
mdp_fused_step_kernel(const float* __restrict__ new_actions, const float* __restrict__ force,
                      float* __restrict__ root_pos_w, float* __restrict__ root_quat_w, int n,
                      const __grid_constant__ RoverMdpParams P, const __grid_constant__ RoverMdpState S,
                      const __grid_constant__ RoverMdpOut O, const __grid_constant__ Tables T,
                      const __grid_constant__ VariatesDev V, long long* __restrict__ out_spawn_index,
                      float* __restrict__ block_stats, unsigned int* __restrict__ done_counter, float* __restrict__ stats,
                      float* __restrict__ log_out, float* __restrict__ obs, int obs_stride, int pre_phases, int phases,
                      const __grid_constant__ StatsExchangeDev X, unsigned long long* __restrict__ lookback) {
    __shared__ int s_bid;
    __shared__ unsigned s_epoch;
    const int n_blocks = (int)gridDim.x;
    int bid = (int)blockIdx.x;
    unsigned epoch = 0u;
    if constexpr (!kRng) {  // (the in-kernel generator needs no reset rank, hence no look-back and no ticket)
        if (threadIdx.x == 0) {
            s_epoch = (unsigned)*reinterpret_cast<volatile unsigned long long*>(lookback + n_blocks + 1);
            s_bid = (int)atomicAdd(lookback + n_blocks, 1ull);
        }
        __syncthreads();
        bid = s_bid;
        epoch = s_epoch;
    }
    const int i = bid * ROVER_MDP_BLOCK + threadIdx.x;
    if (i < n) {  // the post-step's rank-independent inputs: in flight while the pre-step part runs
        asm volatile("prefetch.global.L2 [%0];" ::"l"(root_pos_w + 3 * (size_t)i));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(root_quat_w + 4 * (size_t)i));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(S.pos_cmd_w + 3 * (size_t)i));
        if (!kRng) asm volatile("prefetch.global.L2 [%0];" ::"l"(V.theta_u + (size_t)i * V.n_rounds));
        if ((threadIdx.x & 31) == 0) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(S.heading_cmd_w + i));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(S.time_left + i));
            if (!kRng) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(V.yaw_u + i));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(V.heading_u + i));
            }
            asm volatile("prefetch.global.L2 [%0];" ::"l"(S.err_pos + i));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(S.err_heading + i));
        }
    }
    const bool reset = pre_step_env(i, new_actions, force, n, P, S, O, pre_phases);
    if ((pre_phases & ROVER_PRE_TERMS) && threadIdx.x == 0) O.block_reset_counts[bid] = 0;  // unused by this path
    post_step_block<true, kRng>(bid, n_blocks, reset, root_pos_w, root_quat_w, n, P, S, O, T, V, out_spawn_index, block_stats,
                                done_counter, stats, log_out, obs, obs_stride, phases, X, lookback, epoch);
}

// one block: thread (p, k) = (threadIdx.x / 16, threadIdx.x % 16) reads statistic k of rank p's slot consistently
__global__ void stats_read_kernel(const unsigned char* __restrict__ mailbox, int world, double* __restrict__ out) {
    __shared__ double v[64][kStats];
    const int p = threadIdx.x / kStats, k = threadIdx.x % kStats;
    for (int p0 = 0; p0 < world; p0 += 64) {
        const int pp = p0 + p;
        if (pp < world && p < 64) {
            const unsigned char* slot = mailbox + (size_t)pp * ROVER_MAILBOX_SLOT_BYTES;
            const volatile unsigned long long* sq = reinterpret_cast<const volatile unsigned long long*>(slot);
            const volatile double* val = reinterpret_cast<const volatile double*>(slot + 8);
            double x;
            unsigned long long s1, s2;
            do {  // buffer s1 & 1 holds the totals of sequence s1; it is rewritten only on the way to s1 + 2
                s1 = *sq;
                __threadfence_system();
                x = val[(s1 & 1ull) * kStats + k];
                __threadfence_system();
                s2 = *sq;
            } while (s1 != s2);
            v[p][k] = x;
        }
        __syncthreads();
        if (threadIdx.x < kStats) {
            double t = p0 == 0 ? 0.0 : out[threadIdx.x];
            for (int q = 0; q < min(64, world - p0); ++q) t += v[q][threadIdx.x];
            out[threadIdx.x] = t;
        }
        __syncthreads();
    }
}

}  // namespace rover

static int check_state(const RoverMdpState* s, const RoverMdpOut* o) {
    using namespace rover;
    ROVER_CHECK(s && o, "rover_mdp: NULL state/out struct");
    ROVER_CHECK(s->action && s->prev_action && s->pos_cmd_w && s->heading_cmd_w && s->pos_cmd_b && s->heading_cmd_b &&
                    s->time_left && s->command_counter && s->episode_length_buf && s->episode_sums && s->env_origins &&
                    s->err_pos && s->err_heading,
                "rover_mdp: NULL pointer in RoverMdpState");
    ROVER_CHECK(o->processed_actions && o->joint_pos && o->joint_vel && o->reward && o->term_rewards && o->terminated &&
                        o->truncated && o->term_flags && o->reset_flags && o->block_reset_counts && o->term_values,
                "rover_mdp: NULL pointer in RoverMdpOut");
    ROVER_CHECK((reinterpret_cast<uintptr_t>(o->joint_pos) & 15) == 0, "rover_mdp: joint_pos not 16B aligned");
    return 0;
}

extern "C" int rover_ackermann(const float* actions, int32_t n_envs, const RoverMdpParams* params, float* processed,
                               float* joint_pos, float* joint_vel, void* stream) {
    using namespace rover;
    ROVER_CHECK(n_envs >= 0, "rover_ackermann: negative n_envs");
    if (n_envs == 0) return 0;
    ROVER_CHECK(actions && params && joint_pos && joint_vel, "rover_ackermann: NULL argument");
    ROVER_CHECK(params->action_variant >= 1 && params->action_variant <= 3, "rover_ackermann: action_variant must be 1, 2 or 3");
    ackermann_kernel<<<(n_envs + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(actions, n_envs, *params, processed,
                                                                                        joint_pos, joint_vel);
    return check_launch("ackermann_kernel");
}

extern "C" int rover_mdp_pre_step(const float* new_actions, const float* force_matrix_w, int32_t n_envs,
                                  const RoverMdpParams* params, const RoverMdpState* state, const RoverMdpOut* out,
                                  int32_t phases, void* stream) {
    using namespace rover;
    ROVER_CHECK(n_envs >= 0, "rover_mdp_pre_step: negative n_envs");
    if (n_envs == 0) return 0;
    ROVER_CHECK(params && (new_actions || !(phases & ROVER_PRE_ACTIONS)) &&
                    (force_matrix_w || !(phases & ROVER_PRE_TERMS)),
                "rover_mdp_pre_step: NULL argument");
    ROVER_CHECK(params->num_bodies >= 0 && params->max_episode_length > 0, "rover_mdp_pre_step: bad params");
    if (int rc = check_state(state, out)) return rc;
    const int blocks = (n_envs + ROVER_MDP_BLOCK - 1) / ROVER_MDP_BLOCK;
    mdp_pre_step_kernel<<<blocks, ROVER_MDP_BLOCK, 0, static_cast<cudaStream_t>(stream)>>>(
        new_actions, force_matrix_w, n_envs, *params, *state, *out, phases);
    return check_launch("mdp_pre_step_kernel");
}

namespace rover {

static int make_variates(const RoverResetVariates* var, int n_envs, VariatesDev& V, const char* who) {
    ROVER_CHECK(var != nullptr, "%s: NULL RoverResetVariates", who);
    ROVER_CHECK(var->n_rounds >= 1 && var->n_rounds <= 4096, "%s: n_rounds must be in [1, 4096]", who);
    if (var->rng_state != nullptr) {
        V = VariatesDev{nullptr, nullptr, nullptr, nullptr, reinterpret_cast<unsigned long long*>(var->rng_state), var->n_rounds};
    } else {
        ROVER_CHECK(var->spawn_perm && var->yaw_u && var->heading_u && var->theta_u,
                    "%s: explicit variates need spawn_perm, yaw_u, heading_u and theta_u (or set rng_state)", who);
        V = VariatesDev{reinterpret_cast<const long long*>(var->spawn_perm), var->yaw_u, var->heading_u, var->theta_u, nullptr,
                        var->n_rounds};
    }
    (void)n_envs;
    return 0;
}

static int make_exchange(const RoverStatsExchange* xchg, StatsExchangeDev& X, const char* who) {
    X = StatsExchangeDev{nullptr, nullptr, nullptr, 0, 0};
    if (xchg != nullptr) {
        ROVER_CHECK(xchg->peer_mailbox && xchg->cumulative && xchg->sequence && xchg->world >= 1 && xchg->rank >= 0 &&
                        xchg->rank < xchg->world,
                    "%s: bad RoverStatsExchange", who);
        X = StatsExchangeDev{xchg->peer_mailbox, xchg->cumulative, reinterpret_cast<unsigned long long*>(xchg->sequence),
                             xchg->rank, xchg->world};
    }
    return 0;
}

static int check_post_args(const float* root_pos_w, const float* root_quat_w, int n_envs, const RoverMdpParams* params,
                           const RoverTerrainTables* tables, const float* stats, const float* scratch, const float* obs,
                           int obs_stride, int phases, const char* who) {
    ROVER_CHECK(root_pos_w && root_quat_w && params && tables && stats && scratch, "%s: NULL argument", who);
    ROVER_CHECK(tables->heightmap && tables->safe_mask && tables->spawn_table && tables->height > 0 &&
                    tables->width > 0 && tables->n_spawns > 0 && tables->resolution > 0.f,
                "%s: bad terrain tables", who);
    // without-replacement spawn draw: reset rank j < n_envs indexes a permutation of the table's rows
    ROVER_CHECK(!(phases & ROVER_PHASE_SPAWN) || tables->n_spawns >= n_envs,
                "%s: spawn table has %d rows for %d envs (randperm(len(spawn))[:K] needs K <= len)", who, tables->n_spawns,
                n_envs);
    // a reset resample and a time-based resample of one env share the env's variates; they never meet in one step as
    // long as a fresh command outlives the step (reference: 150 s against 0.2 s)
    ROVER_CHECK(params->resampling_time > params->step_dt, "%s: resampling_time must exceed step_dt", who);
    ROVER_CHECK((reinterpret_cast<uintptr_t>(root_quat_w) & 15) == 0, "%s: root_quat_w not 16B aligned", who);
    ROVER_CHECK(obs == nullptr || obs_stride >= 4, "%s: obs_stride < 4", who);
    return 0;
}

}  // namespace rover

extern "C" int rover_mdp_post_step(float* root_pos_w, float* root_quat_w, int32_t n_envs, const RoverMdpParams* params,
                                   const RoverMdpState* state, const RoverMdpOut* out,
                                   const RoverTerrainTables* tables, const int64_t* spawn_perm, const float* yaw_u,
                                   const float* heading_u, const float* theta_u, int32_t n_rounds,
                                   int64_t* out_spawn_index, float* stats, float* scratch, float* obs,
                                   int32_t obs_stride, int32_t phases, void* stream) {
    return rover_mdp_post_step_x(root_pos_w, root_quat_w, n_envs, params, state, out, tables, spawn_perm, yaw_u, heading_u,
                                 theta_u, n_rounds, out_spawn_index, stats, scratch, obs, obs_stride, phases, nullptr,
                                 stream);
}

extern "C" int rover_mdp_post_step_x(float* root_pos_w, float* root_quat_w, int32_t n_envs, const RoverMdpParams* params,
                                     const RoverMdpState* state, const RoverMdpOut* out,
                                     const RoverTerrainTables* tables, const int64_t* spawn_perm, const float* yaw_u,
                                     const float* heading_u, const float* theta_u, int32_t n_rounds,
                                     int64_t* out_spawn_index, float* stats, float* scratch, float* obs,
                                     int32_t obs_stride, int32_t phases, const RoverStatsExchange* xchg, void* stream) {
    const RoverResetVariates var{spawn_perm, yaw_u, heading_u, theta_u, n_rounds, 0, nullptr};
    return rover_mdp_post_step_v3(root_pos_w, root_quat_w, n_envs, params, state, out, tables, &var, out_spawn_index, stats,
                                  scratch, nullptr, obs, obs_stride, phases, xchg, stream);
}

extern "C" int rover_mdp_post_step_v3(float* root_pos_w, float* root_quat_w, int32_t n_envs, const RoverMdpParams* params,
                                      const RoverMdpState* state, const RoverMdpOut* out,
                                      const RoverTerrainTables* tables, const RoverResetVariates* variates,
                                      int64_t* out_spawn_index, float* stats, float* scratch, float* log_out, float* obs,
                                      int32_t obs_stride, int32_t phases, const RoverStatsExchange* xchg, void* stream) {
    using namespace rover;
    StatsExchangeDev X;
    if (int rc = make_exchange(xchg, X, "rover_mdp_post_step")) return rc;
    ROVER_CHECK(n_envs >= 0, "rover_mdp_post_step: negative n_envs");
    if (n_envs == 0) return 0;
    if (int rc = check_post_args(root_pos_w, root_quat_w, n_envs, params, tables, stats, scratch, obs, obs_stride, phases,
                                 "rover_mdp_post_step"))
        return rc;
    VariatesDev V;
    if (int rc = make_variates(variates, n_envs, V, "rover_mdp_post_step")) return rc;
    if (int rc = check_state(state, out)) return rc;
    Tables T{tables->heightmap, tables->safe_mask, tables->height,   tables->width,   tables->offset_x,
             tables->offset_y,  tables->resolution, tables->spawn_table, tables->n_spawns};
    const int blocks = (n_envs + ROVER_MDP_BLOCK - 1) / ROVER_MDP_BLOCK;
    // scratch layout: [blocks * 16] block partials, then one uint32 completion counter (zero-initialised by caller)
    float* block_stats = scratch;
    unsigned int* counter = reinterpret_cast<unsigned int*>(scratch + (size_t)blocks * ROVER_STATS_LEN);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (V.rng != nullptr)
        mdp_post_step_kernel<true><<<blocks, ROVER_MDP_BLOCK, 0, st>>>(
            root_pos_w, root_quat_w, n_envs, *params, *state, *out, T, V, reinterpret_cast<long long*>(out_spawn_index),
            block_stats, counter, stats, log_out, obs, obs_stride, phases, X);
    else
        mdp_post_step_kernel<false><<<blocks, ROVER_MDP_BLOCK, 0, st>>>(
            root_pos_w, root_quat_w, n_envs, *params, *state, *out, T, V, reinterpret_cast<long long*>(out_spawn_index),
            block_stats, counter, stats, log_out, obs, obs_stride, phases, X);
    return check_launch("mdp_post_step_kernel");
}

extern "C" int rover_mdp_step(const float* new_actions, const float* force_matrix_w, float* root_pos_w, float* root_quat_w,
                              int32_t n_envs, const RoverMdpParams* params, const RoverMdpState* state,
                              const RoverMdpOut* out, const RoverTerrainTables* tables, const int64_t* spawn_perm,
                              const float* yaw_u, const float* heading_u, const float* theta_u, int32_t n_rounds,
                              int64_t* out_spawn_index, float* stats, float* scratch, uint64_t* lookback, float* obs,
                              int32_t obs_stride, int32_t pre_phases, int32_t phases, const RoverStatsExchange* xchg,
                              void* stream) {
    const RoverResetVariates var{spawn_perm, yaw_u, heading_u, theta_u, n_rounds, 0, nullptr};
    return rover_mdp_step_v3(new_actions, force_matrix_w, root_pos_w, root_quat_w, n_envs, params, state, out, tables, &var,
                             out_spawn_index, stats, scratch, lookback, nullptr, obs, obs_stride, pre_phases, phases, xchg,
                             stream);
}

extern "C" int rover_mdp_step_v3(const float* new_actions, const float* force_matrix_w, float* root_pos_w,
                                 float* root_quat_w, int32_t n_envs, const RoverMdpParams* params,
                                 const RoverMdpState* state, const RoverMdpOut* out, const RoverTerrainTables* tables,
                                 const RoverResetVariates* variates, int64_t* out_spawn_index, float* stats, float* scratch,
                                 uint64_t* lookback, float* log_out, float* obs, int32_t obs_stride, int32_t pre_phases,
                                 int32_t phases, const RoverStatsExchange* xchg, void* stream) {
    using namespace rover;
    ROVER_CHECK(n_envs >= 0, "rover_mdp_step: negative n_envs");
    if (n_envs == 0) return 0;
    ROVER_CHECK(params && (new_actions || !(pre_phases & ROVER_PRE_ACTIONS)) &&
                    (force_matrix_w || !(pre_phases & ROVER_PRE_TERMS)) && variates &&
                    (lookback || variates->rng_state),
                "rover_mdp_step: NULL argument (lookback is needed with explicit variates)");
    ROVER_CHECK(params->num_bodies >= 0 && params->max_episode_length > 0, "rover_mdp_step: bad params");
    if (int rc = check_post_args(root_pos_w, root_quat_w, n_envs, params, tables, stats, scratch, obs, obs_stride, phases,
                                 "rover_mdp_step"))
        return rc;
    VariatesDev V;
    if (int rc = make_variates(variates, n_envs, V, "rover_mdp_step")) return rc;
    if (int rc = check_state(state, out)) return rc;
    StatsExchangeDev X;
    if (int rc = make_exchange(xchg, X, "rover_mdp_step")) return rc;
    Tables T{tables->heightmap, tables->safe_mask, tables->height,   tables->width,   tables->offset_x,
             tables->offset_y,  tables->resolution, tables->spawn_table, tables->n_spawns};
    const int blocks = (n_envs + ROVER_MDP_BLOCK - 1) / ROVER_MDP_BLOCK;
    float* block_stats = scratch;
    unsigned int* counter = reinterpret_cast<unsigned int*>(scratch + (size_t)blocks * ROVER_STATS_LEN);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (V.rng != nullptr)
        mdp_fused_step_kernel<true><<<blocks, ROVER_MDP_BLOCK, 0, st>>>(
            new_actions, force_matrix_w, root_pos_w, root_quat_w, n_envs, *params, *state, *out, T, V,
            reinterpret_cast<long long*>(out_spawn_index), block_stats, counter, stats, log_out, obs, obs_stride, pre_phases,
            phases, X, reinterpret_cast<unsigned long long*>(lookback));
    else
        mdp_fused_step_kernel<false><<<blocks, ROVER_MDP_BLOCK, 0, st>>>(
            new_actions, force_matrix_w, root_pos_w, root_quat_w, n_envs, *params, *state, *out, T, V,
            reinterpret_cast<long long*>(out_spawn_index), block_stats, counter, stats, log_out, obs, obs_stride, pre_phases,
            phases, X, reinterpret_cast<unsigned long long*>(lookback));
    return check_launch("mdp_fused_step_kernel");
}

// The variates the kernels generate in registers, evaluated on the HOST by the same functions (rng.cuh): the oracle is
// fed these, the kernel generates its own, and the two must agree bit for bit (tests/test_gpu_rng.py).
extern "C" int rover_rng_variates(uint64_t seed, uint64_t step, int32_t n_envs, int32_t n_rounds, int32_t n_spawns,
                                  int64_t* spawn_by_env, float* yaw_u, float* heading_u, float* theta_u) {
    using namespace rover;
    ROVER_CHECK(n_envs >= 0 && n_rounds >= 1 && n_spawns >= 1, "rover_rng_variates: bad sizes");
    const RngKey key = make_rng_key(seed, step);
    if (spawn_by_env != nullptr) {
        const SpawnPermKey pk = make_spawn_perm_key(key, (uint32_t)n_spawns);
        const int k = n_envs < n_spawns ? n_envs : n_spawns;
        for (int i = 0; i < k; ++i) spawn_by_env[i] = (int64_t)spawn_perm_at(pk, (uint32_t)i);
    }
    for (int i = 0; i < n_envs; ++i) {
        uint32_t w[4];
        rng_env_stream(key, (uint32_t)i, 0u, w);
        if (yaw_u) yaw_u[i] = u01(w[0]);
        if (heading_u) heading_u[i] = u01(w[1]);
        if (theta_u) {
            for (int r0 = 0; r0 < n_rounds; r0 += 4) {
                rng_env_stream(key, (uint32_t)i, 1u + (uint32_t)(r0 / 4), w);
                for (int k = 0; k < 4 && r0 + k < n_rounds; ++k) theta_u[(size_t)i * n_rounds + r0 + k] = u01(w[k]);
            }
        }
    }
    return 0;
}

extern "C" int rover_philox4x32_10(const uint32_t counter[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t o[4];
    rover::philox4x32_10(counter[0], counter[1], counter[2], counter[3], key[0], key[1], o);
    for (int k = 0; k < 4; ++k) out[k] = o[k];
    return 0;
}

extern "C" int rover_stats_read(const void* mailbox_local, int32_t world, double* out, void* stream) {
    using namespace rover;
    ROVER_CHECK(mailbox_local && out && world >= 1, "rover_stats_read: bad arguments");
    stats_read_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const unsigned char*>(mailbox_local), world,
                                                                        out);
    return check_launch("stats_read_kernel");
}

extern "C" int rover_p2p_alloc(void** out_ptr, int64_t bytes) {
    using namespace rover;
    ROVER_CHECK(out_ptr && bytes > 0, "rover_p2p_alloc: bad arguments");
    ROVER_CUDA(cudaMalloc(out_ptr, (size_t)bytes));
    ROVER_CUDA(cudaMemset(*out_ptr, 0, (size_t)bytes));
    ROVER_CUDA(cudaDeviceSynchronize());
    return 0;
}
extern "C" int rover_p2p_free(void* ptr) {
    using namespace rover;
    ROVER_CUDA(cudaFree(ptr));
    return 0;
}
extern "C" int rover_p2p_export(void* ptr, uint8_t handle_out[64]) {
    using namespace rover;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    ROVER_CUDA(cudaIpcGetMemHandle(&h, ptr));
    memcpy(handle_out, &h, 64);
    return 0;
}
extern "C" int rover_p2p_open(const uint8_t handle[64], void** out_ptr) {
    using namespace rover;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    ROVER_CUDA(cudaIpcOpenMemHandle(out_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
extern "C" int rover_p2p_close(void* ptr) {
    using namespace rover;
    ROVER_CUDA(cudaIpcCloseMemHandle(ptr));
    return 0;
}
