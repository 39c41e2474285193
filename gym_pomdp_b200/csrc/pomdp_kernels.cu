// pomdp_kernels.cu -- sm_100a kernels + the C ABI of include/pomdp_b200.h.
//
// Data movement design:
//  * W=1/W=2 envs (Rock, Tag, Tiger, Network): SoA int32 streams.  One thread owns FOUR
//    consecutive env instances: 16-byte vector loads of state/action, 16-byte vector
//    stores of next_state/obs/reward/flags (or next_state + one packed result word),
//    fully coalesced (a warp moves 512 B per instruction).  Persistent grid-stride loop
//    with a balanced trip count, at most SMs x resident CTAs.
//  * Static maps -- Rock's header + transition LUT (17.6 KB for Rock(11,11)), Tag's board
//    tables (4.3 KB) -- are built on the host, owned by the caller and copied global ->
//    shared once per CTA with ONE TMA bulk copy (cp.async.bulk + mbarrier complete_tx);
//    in the step kernel the first global loads are issued before the wait so the table
//    fetch hides under them.  Lookups are per-thread divergent, which is what shared
//    memory (not the constant bank) is for.
//  * BattleShip (W=8, 32 B per board): the step kernel moves board tiles global -> shared
//    and shared -> global with TMA bulk copies (8 KB per 256-env tile); reset is one
//    THREAD per env on 128-bit bitboards (and, as an alternative with identical results,
//    one WARP per env with a ballot scan over the 4*n_tiles placement candidates).
//  * Philox4x32-10 in registers, one block per draw slot per four envs; no RNG state in memory.
//  * policy / rollout / obs_prob / legal_mask / belief_hist: see the comments at each kernel.
//
// There is no CPU path in this file: every entry point launches a kernel.
#include <type_traits>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "pomdp_core.h"
#include "pomdp_envs.h"
#include "pomdp_host.h"

using namespace pomdp;

#define POMDP_THREADS 256
// step kernels: CTA size and the min-resident-CTAs register hint (tuned on B200)
#ifndef POMDP_STEP_THREADS
#define POMDP_STEP_THREADS 512
#endif
#ifndef POMDP_STEP_MINB
#define POMDP_STEP_MINB 2
#endif

// ----------------------------------------------------------------------------- PTX ---
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// TMA 1-D bulk copy shared -> global (bulk async-group completion).
__device__ __forceinline__ void tma_bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
                 "r"(bytes)
                 : "memory");
}
// Programmatic dependent launch (PDL): a step kernel is launched with the programmatic-stream-
// serialization attribute, so its CTAs may be scheduled -- and run their prologue: barrier init,
// TMA copy of the static table -- while the previous kernel in the stream is still draining.
// pdl_wait() blocks until that kernel has completed and its writes are visible; nothing that
// another kernel may have produced is touched before it.  pdl_launch_dependents() lets the NEXT
// kernel's CTAs start filling SM slots as soon as this kernel's CTAs retire.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// Streaming (evict-first) vector accesses: every byte is touched once per launch.
__device__ __forceinline__ int4 ld_stream4(const int32_t* p) { return __ldcs(reinterpret_cast<const int4*>(p)); }
__device__ __forceinline__ void st_stream4(int32_t* p, int4 v) { __stcs(reinterpret_cast<int4*>(p), v); }
__device__ __forceinline__ void st_stream4(float* p, float4 v) { __stcs(reinterpret_cast<float4*>(p), v); }

// Four consecutive packed states (32- or 64-bit each) as 16-byte vectors.
template <typename S> struct StateVec;
template <> struct StateVec<uint32_t> {
    int4 v;
    __device__ __forceinline__ void zero() { v = make_int4(0, 0, 0, 0); }
    __device__ __forceinline__ void load(const int32_t* base, int64_t i) { v = ld_stream4(base + i); }
    __device__ __forceinline__ void unpack(uint32_t s[4]) const {
        s[0] = (uint32_t)v.x; s[1] = (uint32_t)v.y; s[2] = (uint32_t)v.z; s[3] = (uint32_t)v.w;
    }
    __device__ static __forceinline__ void store(int32_t* base, int64_t i, const uint32_t s[4]) {
        st_stream4(base + i, make_int4((int)s[0], (int)s[1], (int)s[2], (int)s[3]));
    }
};
// 256-bit global accesses (sm_100: LDG.E.256 / STG.E.256): four 64-bit packed states in ONE instruction
__device__ __forceinline__ void ld_stream_4x64(const void* p, uint64_t s[4]) {
    asm volatile("ld.global.cs.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(s[0]), "=l"(s[1]), "=l"(s[2]), "=l"(s[3]) : "l"(p));
}
__device__ __forceinline__ void st_stream_4x64(void* p, const uint64_t s[4]) {
    asm volatile("st.global.cs.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(p), "l"(s[0]), "l"(s[1]), "l"(s[2]), "l"(s[3]) : "memory");
}
template <> struct StateVec<uint64_t> {
    uint64_t v[4];
    __device__ __forceinline__ void zero() { v[0] = v[1] = v[2] = v[3] = 0; }
    __device__ __forceinline__ void load(const int32_t* base, int64_t i) {
        const int32_t* p = base + 2 * i;
        if ((reinterpret_cast<uintptr_t>(base) & 31) == 0) {      // uniform: the whole array is 32-byte aligned
            ld_stream_4x64(p, v);
        } else {
            const int4 a = ld_stream4(p), b = ld_stream4(p + 4);
            v[0] = (uint32_t)a.x | ((uint64_t)(uint32_t)a.y << 32); v[1] = (uint32_t)a.z | ((uint64_t)(uint32_t)a.w << 32);
            v[2] = (uint32_t)b.x | ((uint64_t)(uint32_t)b.y << 32); v[3] = (uint32_t)b.z | ((uint64_t)(uint32_t)b.w << 32);
        }
    }
    __device__ __forceinline__ void unpack(uint64_t s[4]) const { s[0] = v[0]; s[1] = v[1]; s[2] = v[2]; s[3] = v[3]; }
    __device__ static __forceinline__ void store(int32_t* base, int64_t i, const uint64_t s[4]) {
        int32_t* p = base + 2 * i;
        if ((reinterpret_cast<uintptr_t>(base) & 31) == 0) {
            st_stream_4x64(p, s);
        } else {
            st_stream4(p, make_int4((int)(uint32_t)s[0], (int)(s[0] >> 32), (int)(uint32_t)s[1], (int)(s[1] >> 32)));
            st_stream4(p + 4, make_int4((int)(uint32_t)s[2], (int)(s[2] >> 32), (int)(uint32_t)s[3], (int)(s[3] >> 32)));
        }
    }
};
__device__ __forceinline__ uint32_t load_state1(const int32_t* base, int64_t i, uint32_t) { return (uint32_t)base[i]; }
__device__ __forceinline__ uint64_t load_state1(const int32_t* base, int64_t i, uint64_t) {
    return (uint32_t)base[2 * i] | ((uint64_t)(uint32_t)base[2 * i + 1] << 32);
}
__device__ __forceinline__ void store_state1(int32_t* base, int64_t i, uint32_t s) { base[i] = (int32_t)s; }
__device__ __forceinline__ void store_state1(int32_t* base, int64_t i, uint64_t s) {
    base[2 * i] = (int32_t)(uint32_t)s;
    base[2 * i + 1] = (int32_t)(s >> 32);
}

// Envs whose static tables travel inside the kernel parameters (Network: 2.8 KB) copy them to shared memory once per
// CTA: the lookups are per-thread divergent, which the constant bank would serialise.
template <class Env>
__device__ __forceinline__ void stage_param_table(unsigned char* smem_table, const typename Env::Params& p) {
    if constexpr (Env::kParamTable) {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(Env::param_table(p));
        uint32_t* dst = reinterpret_cast<uint32_t*>(smem_table);
        for (uint32_t i = threadIdx.x; i < Env::kParamTableBytes / 4; i += blockDim.x) dst[i] = src[i];
        __syncthreads();
    }
}

// ----------------------------------------------------------------- step (streams) ---
// kVec: all six arrays are 16-byte aligned and global_offset is a multiple of 4 -> every
// thread owns aligned groups of FOUR envs (16-byte vector loads/stores, one Philox call per
// draw slot per group).  The loop is software-pipelined: the loads of a thread's next group
// are issued before the current group is computed, so two groups' worth of bytes per
// thread are in flight.  Otherwise: scalar thread-per-env path (oddly offset views).
// kPacked: obs / reward / flags leave as ONE int32 stream (pack_result, pomdp_core.h) written to `obs`; `reward` and
// `flags` are unused.  16 instead of 24 bytes per env-step on the device, 8 instead of 16 for a host caller to fetch.
template <class Env, bool kVec, bool kPacked = false>
__global__ void __launch_bounds__(POMDP_STEP_THREADS, POMDP_STEP_MINB)
pomdp_step_kernel(const __grid_constant__ typename Env::Params p, const void* __restrict__ g_table,
                  const int32_t* state, const int32_t* __restrict__ action, int32_t* next_state,
                  int32_t* __restrict__ obs, float* __restrict__ reward, int32_t* __restrict__ flags, int64_t n,
                  uint64_t goff, const __grid_constant__ PhiloxKey seed, uint32_t step_ctr, uint32_t table_bytes) {
    typedef typename Env::State S;
    extern __shared__ __align__(128) unsigned char smem_table[];
    __shared__ alignas(8) uint64_t bar;
    const unsigned char* lut = smem_table;
    if (Env::kTable) {
        if (threadIdx.x == 0) {
            mbar_init(&bar, 1);
            fence_mbar_init();
            mbar_expect_tx(&bar, table_bytes);
            tma_bulk_g2s(smem_table, g_table, table_bytes, &bar);
        }
        __syncthreads();   // barrier object initialised before anyone polls it
    }
    stage_param_table<Env>(smem_table, p);

    pdl_wait();                // everything above overlapped the previous kernel's tail; state/action may be its outputs
    pdl_launch_dependents();
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    bool table_ready = !Env::kTable;
    if (kVec) {
        const int64_t n_groups = n >> 2;
        const uint64_t group0 = goff >> 2;
        int64_t g = tid;
        StateVec<S> cur_s;
        int4 cur_a = make_int4(0, 0, 0, 0);
        if (g < n_groups) {
            cur_s.load(state, g << 2);
            cur_a = ld_stream4(action + (g << 2));
        }
        if (!table_ready) { mbar_wait(&bar, 0); table_ready = true; }   // the loads above are already in flight
        while (g < n_groups) {
            const int64_t gn = g + nthreads;
            StateVec<S> nxt_s = cur_s;
            int4 nxt_a = cur_a;
            if (gn < n_groups) {
                nxt_s.load(state, gn << 2);
                nxt_a = ld_stream4(action + (gn << 2));
            }
            const int64_t i = g << 2;
            S s[4], s2[4];
            cur_s.unpack(s);
            const int32_t a[4] = {cur_a.x, cur_a.y, cur_a.z, cur_a.w};
            int32_t ob[4], fl[4];
            float rw[4];
            Env::step4(p, lut, s, a, seed, group0 + (uint64_t)g, step_ctr, s2, ob, rw, fl);
            StateVec<S>::store(next_state, i, s2);
            if (kPacked) {
                st_stream4(obs + i, make_int4(pack_result(ob[0], Env::reward_units(rw[0]), fl[0]),
                                              pack_result(ob[1], Env::reward_units(rw[1]), fl[1]),
                                              pack_result(ob[2], Env::reward_units(rw[2]), fl[2]),
                                              pack_result(ob[3], Env::reward_units(rw[3]), fl[3])));
            } else {
                st_stream4(obs + i, make_int4(ob[0], ob[1], ob[2], ob[3]));
                st_stream4(reward + i, make_float4(rw[0], rw[1], rw[2], rw[3]));
                st_stream4(flags + i, make_int4(fl[0], fl[1], fl[2], fl[3]));
            }
            cur_s = nxt_s;
            cur_a = nxt_a;
            g = gn;
        }
        // tail (n % 4 envs): the first few threads of the grid
        const int64_t i = (n_groups << 2) + tid;
        if (i < n) {
            S s2; int32_t ob, fl; float rw;
            Env::step1(p, lut, load_state1(state, i, S()), action[i], seed, goff + (uint64_t)i, step_ctr, s2, ob, rw, fl);
            store_state1(next_state, i, s2);
            if (kPacked) obs[i] = pack_result(ob, Env::reward_units(rw), fl);
            else { obs[i] = ob; reward[i] = rw; flags[i] = fl; }
        }
    } else {
        for (int64_t i = tid; i < n; i += nthreads) {
            const S s = load_state1(state, i, S());
            const int32_t a = action[i];
            if (!table_ready) { mbar_wait(&bar, 0); table_ready = true; }
            S s2; int32_t ob, fl; float rw;
            Env::step1(p, lut, s, a, seed, goff + (uint64_t)i, step_ctr, s2, ob, rw, fl);
            store_state1(next_state, i, s2);
            if (kPacked) obs[i] = pack_result(ob, Env::reward_units(rw), fl);
            else { obs[i] = ob; reward[i] = rw; flags[i] = fl; }
        }
    }
    // a CTA must not exit while its bulk copy may still be in flight
    if (!table_ready) mbar_wait(&bar, 0);
}

// ------------------------------------------------------------- diagnostics: stream probe ---
// The step kernels' memory behaviour and nothing else: per group of four envs two 16-byte streaming loads (state,
// action) and four 16-byte streaming stores, the same persistent grid, software pipelining and PDL -- no table, no
// Philox, no transition.  bench.py times it next to the real step: what this 1:2 read:write mix over six streams reaches
// on the part is the roof the RockSample step is measured against.  W = 1 layouts only.
__global__ void __launch_bounds__(POMDP_STEP_THREADS, POMDP_STEP_MINB)
pomdp_stream_probe_kernel(const int32_t* state, const int32_t* __restrict__ action, int32_t* next_state,
                          int32_t* __restrict__ obs, float* __restrict__ reward, int32_t* __restrict__ flags, int64_t n) {
    pdl_wait();
    pdl_launch_dependents();
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    const int64_t n_groups = n >> 2;
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int4 cs = make_int4(0, 0, 0, 0), ca = cs;
    if (g < n_groups) { cs = ld_stream4(state + (g << 2)); ca = ld_stream4(action + (g << 2)); }
    while (g < n_groups) {
        const int64_t gn = g + nthreads;
        int4 ns = cs, na = ca;
        if (gn < n_groups) { ns = ld_stream4(state + (gn << 2)); na = ld_stream4(action + (gn << 2)); }
        const int4 v = make_int4(cs.x ^ ca.x, cs.y ^ ca.y, cs.z ^ ca.z, cs.w ^ ca.w);
        const int64_t i = g << 2;
        st_stream4(next_state + i, v);
        st_stream4(obs + i, ca);
        st_stream4(reward + i, make_float4(__int_as_float(cs.x), __int_as_float(cs.y), __int_as_float(cs.z), __int_as_float(cs.w)));
        st_stream4(flags + i, v);
        cs = ns; ca = na; g = gn;
    }
}

// The same probe for two-word states (Rock(15,15)): 32 bytes of state per group move as ONE 256-bit access each way, like
// StateVec<uint64_t> does in the step kernel; 12 bytes read and 20 written per env.
__global__ void __launch_bounds__(POMDP_STEP_THREADS, POMDP_STEP_MINB)
pomdp_stream_probe2_kernel(const int32_t* state, const int32_t* __restrict__ action, int32_t* next_state,
                           int32_t* __restrict__ obs, float* __restrict__ reward, int32_t* __restrict__ flags, int64_t n) {
    pdl_wait();
    pdl_launch_dependents();
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    const int64_t n_groups = n >> 2;
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    StateVec<uint64_t> cs;
    cs.zero();
    int4 ca = make_int4(0, 0, 0, 0);
    if (g < n_groups) { cs.load(state, g << 2); ca = ld_stream4(action + (g << 2)); }
    while (g < n_groups) {
        const int64_t gn = g + nthreads;
        StateVec<uint64_t> ns = cs;
        int4 na = ca;
        if (gn < n_groups) { ns.load(state, gn << 2); na = ld_stream4(action + (gn << 2)); }
        uint64_t s[4];
        cs.unpack(s);
        const uint64_t s2[4] = {s[0] ^ (uint32_t)ca.x, s[1] ^ (uint32_t)ca.y, s[2] ^ (uint32_t)ca.z, s[3] ^ (uint32_t)ca.w};
        const int64_t i = g << 2;
        StateVec<uint64_t>::store(next_state, i, s2);
        st_stream4(obs + i, ca);
        st_stream4(reward + i, make_float4(__int_as_float((int)s[0]), __int_as_float((int)s[1]), __int_as_float((int)s[2]), __int_as_float((int)s[3])));
        st_stream4(flags + i, make_int4((int)(s[0] >> 32), (int)(s[1] >> 32), (int)(s[2] >> 32), (int)(s[3] >> 32)));
        cs = ns; ca = na; g = gn;
    }
}

// ---------------------------------------------------------------- reset (streams) ---
// kVec (state 16-byte aligned, global_offset % 4 == 0): four envs per thread, one Philox call
// per draw slot per group, vector stores when the whole group is reset.
// kFast: the common call -- no mask, obs present and 16-byte aligned: the loop body is reset4 + two 16-byte streaming
// stores and nothing else (the masked variant carries four byte loads and per-element predicates through the loop,
// which costs more issue slots than the Philox call itself in a kernel that only writes 8 bytes per env).
template <class Env, bool kVec, bool kFast = false>
__global__ void __launch_bounds__(POMDP_THREADS)
pomdp_reset_kernel(const __grid_constant__ typename Env::Params p, int32_t* __restrict__ state,
                   int32_t* __restrict__ obs, const uint8_t* __restrict__ mask, int64_t n, uint64_t goff,
                   const __grid_constant__ PhiloxKey seed, uint32_t step_ctr) {
    typedef typename Env::State S;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    int64_t scalar_from = 0;
    if (kVec && kFast) {
        const int64_t n_groups = n >> 2;
        const uint64_t group0 = goff >> 2;
        for (int64_t g = tid; g < n_groups; g += nthreads) {
            S s[4];
            int32_t ob[4];
            Env::reset4(p, seed, group0 + (uint64_t)g, step_ctr, s, ob);
            StateVec<S>::store(state, g << 2, s);
            st_stream4(obs + (g << 2), make_int4(ob[0], ob[1], ob[2], ob[3]));
        }
        scalar_from = n_groups << 2;
    } else if (kVec) {
        const int64_t n_groups = n >> 2;
        const bool obs_vec = obs && ((reinterpret_cast<uintptr_t>(obs) & 15) == 0);
        for (int64_t g = tid; g < n_groups; g += nthreads) {
            const int64_t i = g << 2;
            bool m[4] = {true, true, true, true};
            if (mask) {
#pragma unroll
                for (int j = 0; j < 4; ++j) m[j] = mask[i + j] != 0;
            }
            if (!(m[0] || m[1] || m[2] || m[3])) continue;
            S s[4];
            int32_t ob[4];
            Env::reset4(p, seed, (goff >> 2) + (uint64_t)g, step_ctr, s, ob);
            if (m[0] && m[1] && m[2] && m[3]) {
                StateVec<S>::store(state, i, s);
                if (obs_vec) st_stream4(obs + i, make_int4(ob[0], ob[1], ob[2], ob[3]));
                else if (obs) { obs[i] = ob[0]; obs[i + 1] = ob[1]; obs[i + 2] = ob[2]; obs[i + 3] = ob[3]; }
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (m[j]) { store_state1(state, i + j, s[j]); if (obs) obs[i + j] = ob[j]; }
            }
        }
        scalar_from = n_groups << 2;
    }
    for (int64_t i = scalar_from + tid; i < n; i += nthreads) {
        if (mask && !mask[i]) continue;
        S s; int32_t ob;
        Env::reset1(p, seed, goff + (uint64_t)i, step_ctr, s, ob);
        store_state1(state, i, s);
        if (obs) obs[i] = ob;
    }
}

// ------------------------------------------------------- policy / rollout (streams) ---
// Stages the static table (Rock) into shared memory with one TMA bulk copy and waits for it.
template <class Env>
__device__ __forceinline__ void stage_table_sync(unsigned char* smem_table, const void* g_table, uint32_t table_bytes,
                                                 uint64_t* bar) {
    if (Env::kTable) {
        if (threadIdx.x == 0) {
            mbar_init(bar, 1);
            fence_mbar_init();
            mbar_expect_tx(bar, table_bytes);
            tma_bulk_g2s(smem_table, g_table, table_bytes, bar);
        }
        __syncthreads();
        mbar_wait(bar, 0);
    }
}

// action[i] = np.random.choice(env._generate_legal()) for every env: draw (domain POLICY, slot 0) of step_ctr.
template <class Env, bool kVec>
__global__ void __launch_bounds__(POMDP_THREADS)
pomdp_policy_kernel(const __grid_constant__ typename Env::Params p, const void* __restrict__ g_table,
                    const int32_t* __restrict__ state, int32_t* __restrict__ action, int64_t n, uint64_t goff,
                    const __grid_constant__ PhiloxKey seed, uint32_t step_ctr, uint32_t table_bytes) {
    typedef typename Env::State S;
    extern __shared__ __align__(128) unsigned char smem_table[];
    __shared__ alignas(8) uint64_t bar;
    stage_table_sync<Env>(smem_table, g_table, table_bytes, &bar);
    const unsigned char* tbl = smem_table;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    int64_t scalar_from = 0;
    if (kVec) {
        const int64_t n_groups = n >> 2;
        for (int64_t g = tid; g < n_groups; g += nthreads) {
            StateVec<S> v;
            v.load(state, g << 2);
            S s[4];
            v.unpack(s);
            const U4 q = draw_quad(seed, (goff >> 2) + (uint64_t)g, step_ctr, DOMAIN_POLICY, 0);
            st_stream4(action + (g << 2), make_int4(Env::policy(p, tbl, s[0], q.x), Env::policy(p, tbl, s[1], q.y),
                                                    Env::policy(p, tbl, s[2], q.z), Env::policy(p, tbl, s[3], q.w)));
        }
        scalar_from = n_groups << 2;
    }
    for (int64_t i = scalar_from + tid; i < n; i += nthreads)
        action[i] = Env::policy(p, tbl, load_state1(state, i, S()), draw_word(seed, goff + (uint64_t)i, step_ctr, DOMAIN_POLICY, 0));
}

// An env may give the rollouts their own flavour of step4 (same results; Tag: pomdp_envs.h)
template <class Env, class = void> struct HasRolloutStep : std::false_type {};
template <class Env> struct HasRolloutStep<Env, std::void_t<decltype(&Env::step4_rollout)>> : std::true_type {};

// Fused T-step rollout under the uniform-legal policy: the packed states stay in registers for the whole
// rollout; per env the kernel reads 4W bytes and writes 4W + 16 (final state, float64 return, steps, flags).
// One thread owns an aligned group of four envs, so every Philox call (one policy word + the step's own
// slots per time step) serves four envs; all four share the discount gamma^t.
template <class Env, bool kVec>
__global__ void __launch_bounds__(POMDP_THREADS)
pomdp_rollout_kernel(const __grid_constant__ typename Env::Params p, const void* __restrict__ g_table,
                     const int32_t* state, const int32_t* __restrict__ first_action, int32_t* final_state,
                     double* __restrict__ ret, int32_t* __restrict__ steps,
                     int32_t* __restrict__ flags, int64_t n, uint64_t goff, const __grid_constant__ PhiloxKey seed,
                     uint32_t ctr0, int32_t max_steps, double gamma, uint32_t table_bytes) {
    typedef typename Env::State S;
    extern __shared__ __align__(128) unsigned char smem_table[];
    __shared__ alignas(8) uint64_t bar;
    stage_table_sync<Env>(smem_table, g_table, table_bytes, &bar);
    stage_param_table<Env>(smem_table, p);
    const unsigned char* tbl = smem_table;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    int64_t scalar_from = 0;
    if (kVec) {
        const int64_t n_groups = n >> 2;
        for (int64_t g = tid; g < n_groups; g += nthreads) {
            const int64_t i = g << 2;
            const uint64_t group = (goff >> 2) + (uint64_t)g;
            StateVec<S> v;
            v.load(state, i);
            S s[4];
            v.unpack(s);
            double r[4] = {0.0, 0.0, 0.0, 0.0}, disc = 1.0;
            int32_t nst[4] = {0, 0, 0, 0}, facc[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) facc[j] = Env::is_done(s[j]) ? (int32_t)FLAG_DONE : 0;
            int4 fa = make_int4(0, 0, 0, 0);
            if (first_action) fa = ((reinterpret_cast<uintptr_t>(first_action) & 15) == 0)
                                       ? ld_stream4(first_action + i)
                                       : make_int4(first_action[i], first_action[i + 1], first_action[i + 2], first_action[i + 3]);
            for (int32_t t = 0; t < max_steps; ++t) {
                bool act[4], any = false;
#pragma unroll
                for (int j = 0; j < 4; ++j) { act[j] = !Env::is_done(s[j]); any = any || act[j]; }
                if (!any) break;
                const uint32_t ctr = ctr0 + (uint32_t)t;
                const U4 q = draw_quad(seed, group, ctr, DOMAIN_POLICY, 0);
                int32_t a[4] = {Env::policy(p, tbl, s[0], q.x), Env::policy(p, tbl, s[1], q.y),
                                Env::policy(p, tbl, s[2], q.z), Env::policy(p, tbl, s[3], q.w)};
                if (t == 0 && first_action) { a[0] = fa.x; a[1] = fa.y; a[2] = fa.z; a[3] = fa.w; }   // Q(s, a): the caller's action first
                S s2[4];
                int32_t ob[4], fl[4];
                float rw[4];
                if constexpr (HasRolloutStep<Env>::value) Env::step4_rollout(p, tbl, s, a, seed, group, ctr, s2, ob, rw, fl);
                else Env::step4(p, tbl, s, a, seed, group, ctr, s2, ob, rw, fl);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (act[j]) {
                        s[j] = s2[j];
                        r[j] = dadd_rn(r[j], dmul_rn(Env::reward64(rw[j]), disc));
                        ++nst[j];
                        facc[j] |= fl[j];
                    }
                disc = dmul_rn(disc, gamma);
            }
            if (final_state) StateVec<S>::store(final_state, i, s);
            __stcs(reinterpret_cast<double2*>(ret + i), make_double2(r[0], r[1]));
            __stcs(reinterpret_cast<double2*>(ret + i) + 1, make_double2(r[2], r[3]));
            st_stream4(steps + i, make_int4(nst[0], nst[1], nst[2], nst[3]));
            st_stream4(flags + i, make_int4(facc[0], facc[1], facc[2], facc[3]));
        }
        scalar_from = n_groups << 2;
    }
    for (int64_t i = scalar_from + tid; i < n; i += nthreads) {
        S s = load_state1(state, i, S());
        RolloutAcc acc;
        rollout1<Env>(p, tbl, s, seed, goff + (uint64_t)i, ctr0, max_steps, gamma, acc, first_action != nullptr,
                      first_action ? first_action[i] : 0);
        if (final_state) store_state1(final_state, i, s);
        ret[i] = acc.ret; steps[i] = acc.steps; flags[i] = acc.flags;
    }
}

// BattleShip: one thread per board (8 words in registers); the step draws nothing, the policy one word per step.
__device__ __forceinline__ void ship_load(const int32_t* state, int64_t i, bool vec, uint32_t w[SHIP_WORDS]) {
    if (vec) {
        const uint4 lo = __ldcs(reinterpret_cast<const uint4*>(state + i * SHIP_WORDS));
        const uint4 hi = __ldcs(reinterpret_cast<const uint4*>(state + i * SHIP_WORDS) + 1);
        w[0] = lo.x; w[1] = lo.y; w[2] = lo.z; w[3] = lo.w; w[4] = hi.x; w[5] = hi.y; w[6] = hi.z; w[7] = hi.w;
    } else {
#pragma unroll
        for (int k = 0; k < SHIP_WORDS; ++k) w[k] = (uint32_t)state[i * SHIP_WORDS + k];
    }
}
// prob[i] = env._compute_prob(action[i], next_state[i], obs[i]) (float64);  mask[i, :] = env._generate_legal() as bits
template <class Env>
__global__ void __launch_bounds__(POMDP_THREADS)
pomdp_obs_prob_kernel(const __grid_constant__ typename Env::Params p, const void* __restrict__ g_table,
                      const int32_t* __restrict__ state, const int32_t* __restrict__ action, const int32_t* __restrict__ obs,
                      double* __restrict__ prob, int64_t n, double extra, uint32_t table_bytes) {
    typedef typename Env::State S;
    extern __shared__ __align__(128) unsigned char smem_table[];
    __shared__ alignas(8) uint64_t bar;
    stage_table_sync<Env>(smem_table, g_table, table_bytes, &bar);
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nthreads)
        prob[i] = Env::obs_prob(p, smem_table, load_state1(state, i, S()), __ldcs(action + i), __ldcs(obs + i), extra);
}
template <class Env>
__global__ void __launch_bounds__(POMDP_THREADS)
pomdp_legal_mask_kernel(const __grid_constant__ typename Env::Params p, const void* __restrict__ g_table,
                        const int32_t* __restrict__ state, uint32_t* __restrict__ mask, int64_t n, uint32_t table_bytes) {
    typedef typename Env::State S;
    extern __shared__ __align__(128) unsigned char smem_table[];
    __shared__ alignas(8) uint64_t bar;
    stage_table_sync<Env>(smem_table, g_table, table_bytes, &bar);
    const int words = Env::mask_words(p);
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nthreads) {
        uint32_t m[2] = {0u, 0u};
        Env::legal_mask(p, smem_table, load_state1(state, i, S()), m);
        mask[i * words] = m[0];
        if (words > 1) mask[i * words + 1] = m[1];
    }
}
// rock.py:273-291 in the reference's list order (pomdp_core.h: rock_legal_list)
template <typename S>
__global__ void __launch_bounds__(POMDP_THREADS)
pomdp_rock_legal_list_kernel(const __grid_constant__ RockDev p, const void* __restrict__ g_table, const int32_t* __restrict__ state,
                             uint32_t* __restrict__ list, int64_t n, uint32_t table_bytes) {
    extern __shared__ __align__(128) unsigned char smem_table[];
    __shared__ alignas(8) uint64_t bar;
    stage_table_sync<RockEnvT<S, false>>(smem_table, g_table, table_bytes, &bar);
    const RockLut* lut = reinterpret_cast<const RockLut*>(smem_table + ROCK_LUT_OFFSET);
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nthreads)
        list[i] = rock_legal_list<S>(p, lut, load_state1(state, i, S()));
}
// Rock belief side-statistics (rock.py:177-191): one thread per env touches the one rock its check action read.
template <typename S>
__global__ void __launch_bounds__(POMDP_THREADS)
pomdp_rock_belief_update_kernel(const __grid_constant__ RockDev p, const void* __restrict__ g_table,
                                const int32_t* __restrict__ state, const int32_t* __restrict__ action,
                                const int32_t* __restrict__ obs, int32_t* __restrict__ count, int32_t* __restrict__ measured,
                                double* __restrict__ lkv, double* __restrict__ lkw, double* __restrict__ pv, int64_t n,
                                uint32_t table_bytes) {
    extern __shared__ __align__(128) unsigned char smem_table[];
    __shared__ alignas(8) uint64_t bar;
    stage_table_sync<RockEnvT<S, false>>(smem_table, g_table, table_bytes, &bar);
    const RockTableHdr* hdr = reinterpret_cast<const RockTableHdr*>(smem_table);
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nthreads) {
        const int32_t a = __ldcs(action + i), ob = __ldcs(obs + i);
        if (a <= 4 || a >= (int32_t)p.n_actions || ob == 0) continue;
        const int64_t j = i * p.k + (a - 5);
        int32_t c = count[j], m = measured[j];
        double v = lkv[j], w = lkw[j], q = pv[j];
        rock_belief_update<S>(p, hdr, load_state1(state, i, S()), a, ob, c, m, v, w, q);
        count[j] = c; measured[j] = m; lkv[j] = v; lkw[j] = w; pv[j] = q;
    }
}

// ------------------------------------------------ heuristic policies ---
// RockEnv._generate_preferred (rock.py:293-374) on the per-rock planes: belief side-statistics (count, measured,
// prob_valuable: the env's own state, pomdp_rock_belief_update) and the history totals (check_totals:
// pomdp_rock_history_update).  A NULL plane stands for its fresh value (Rock.__init__ rock.py:78-86 / an empty history).
// kPolicy: action[i] = np.random.choice(_generate_preferred(history)) (draw: domain POLICY, slot 0, as the uniform-legal
// policy); else out[i] = the preferred set as a bit mask over action ids, 0 = "fall back to _generate_legal()".
template <typename S, bool kPolicy>
__global__ void __launch_bounds__(POMDP_THREADS)
pomdp_rock_preferred_kernel(const __grid_constant__ RockDev p, const void* __restrict__ g_table, const int32_t* __restrict__ state,
                            const int32_t* __restrict__ count, const int32_t* __restrict__ measured, const double* __restrict__ pv,
                            const int32_t* __restrict__ totals, int32_t* __restrict__ out, int64_t n, uint64_t goff,
                            const __grid_constant__ PhiloxKey seed, uint32_t step_ctr, uint32_t table_bytes) {
    extern __shared__ __align__(128) unsigned char smem_table[];
    __shared__ alignas(8) uint64_t bar;
    stage_table_sync<RockEnvT<S, false>>(smem_table, g_table, table_bytes, &bar);
    const RockTableHdr* hdr = reinterpret_cast<const RockTableHdr*>(smem_table);
    const RockLut* lut = reinterpret_cast<const RockLut*>(smem_table + ROCK_LUT_OFFSET);
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nthreads) {
        const RockPlanesView h = {count, measured, pv, totals, i * p.k};
        const S s = load_state1(state, i, S());
        if (kPolicy) out[i] = rock_policy_preferred<S>(p, hdr, lut, s, h, draw_word(seed, goff + (uint64_t)i, step_ctr, DOMAIN_POLICY, 0));
        else out[i] = (int32_t)rock_preferred_mask<S>(p, hdr, s, h);
    }
}
// One transition appended to every env's history (rock.py:302-309, 325-331): the checked rock's two totals.
__global__ void __launch_bounds__(POMDP_THREADS)
pomdp_rock_history_update_kernel(int k, const int32_t* __restrict__ obs_field, const int32_t* __restrict__ action,
                                 const int32_t* __restrict__ next_obs_field, int32_t* __restrict__ totals, int64_t n) {
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nthreads) {
        const int32_t a = __ldcs(action + i);
        if (a < 5 || a >= 5 + k) continue;
        const int64_t j = i * k + (a - 5);
        const int32_t t = totals[j];
        int32_t ts = rock_totals_sample(t), td = rock_totals_dir(t);
        rock_history_update(a, __ldcs(obs_field + i), __ldcs(next_obs_field + i), ts, td);
        totals[j] = rock_totals_pack(ts, td);
    }
}
// The reference's heuristic rollout loop (rock.py:557-572 with use_heuristic=True) fused into one kernel, one thread per
// env: a = choice(_generate_preferred(history)); next_ob, rw, done = step(a); history.append(Transition(...)); ob = next_ob;
// r += rw * disc; disc *= gamma.  The per-rock planes live in local memory for the whole rollout (k <= 16 rocks x 40 B);
// planes that were passed in are updated in place at the end.  next_is_reward: the transition's `next_observation`
// field holds the reward (the reference's own positional Transition(ob, action, next_ob, rw, done), rock.py:566),
// otherwise the next observation.
template <typename S, bool STOCH, bool kRecords>
__global__ void __launch_bounds__(POMDP_THREADS)
pomdp_rock_rollout_preferred_kernel(const __grid_constant__ RockDev p, const void* __restrict__ g_table, const int32_t* state,
                                    const int32_t* __restrict__ first_action, const __grid_constant__ RockPlanesPtr pl,
                                    int32_t* final_state, double* __restrict__ ret, int32_t* __restrict__ steps,
                                    int32_t* __restrict__ flags, int64_t n, uint64_t goff, const __grid_constant__ PhiloxKey seed,
                                    uint32_t ctr0, int32_t max_steps, double gamma, int32_t next_is_reward, uint32_t table_bytes) {
    typedef RockEnvT<S, STOCH> Env;
    extern __shared__ __align__(128) unsigned char smem_table[];
    __shared__ alignas(8) uint64_t bar;
    stage_table_sync<Env>(smem_table, g_table, table_bytes, &bar);
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nthreads) {
        const int64_t base = i * p.k;
        int32_t prev_ob = pl.prev_obs ? pl.prev_obs[i] : 0;                     // RockEnv.reset returns Obs.NULL (rock.py:241)
        S s = load_state1(state, i, S());
        RolloutAcc acc;
        if (kRecords) {         // fresh planes in, none out: one lazily touched 32-byte record per rock in the caller's scratch
            RockHeurRecords h;
            h.init(reinterpret_cast<RockRec*>(pl.scratch) + base);
            rock_rollout_preferred1<S, STOCH>(p, smem_table, s, seed, goff + (uint64_t)i, ctr0, max_steps, gamma, next_is_reward != 0,
                                              first_action != nullptr, first_action ? first_action[i] : 0, h, prev_ob, acc);
        } else {
            RockHeurLocal h;
            h.load(pl, base, p.k);
            rock_rollout_preferred1<S, STOCH>(p, smem_table, s, seed, goff + (uint64_t)i, ctr0, max_steps, gamma, next_is_reward != 0,
                                              first_action != nullptr, first_action ? first_action[i] : 0, h, prev_ob, acc);
            h.store(pl, base, p.k);
        }
        if (final_state) store_state1(final_state, i, s);
        ret[i] = acc.ret; steps[i] = acc.steps; flags[i] = acc.flags;
        if (pl.prev_obs) pl.prev_obs[i] = prev_ob;
    }
}

// TagEnv._generate_preferred (tag.py:231-243): needs only the last (observation, action) of the history.
template <bool kPolicy>
__global__ void __launch_bounds__(POMDP_THREADS)
pomdp_tag_preferred_kernel(const void* __restrict__ g_table, const int32_t* __restrict__ state, const int32_t* __restrict__ last_obs,
                           const int32_t* __restrict__ last_action, int32_t* __restrict__ out, int64_t n, uint64_t goff,
                           const __grid_constant__ PhiloxKey seed, uint32_t step_ctr) {
    extern __shared__ __align__(128) unsigned char smem_table[];
    __shared__ alignas(8) uint64_t bar;
    stage_table_sync<TagEnvT<1>>(smem_table, g_table, TAG_TABLES_BASE_BYTES, &bar);
    const TagTables* T = reinterpret_cast<const TagTables*>(smem_table);
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nthreads) {
        const uint32_t s = (uint32_t)state[i];
        const int32_t lo = last_obs ? last_obs[i] : 0, la = last_action ? last_action[i] : -1;
        if (kPolicy) out[i] = tag_policy_preferred(T, s, lo, la, draw_word(seed, goff + (uint64_t)i, step_ctr, DOMAIN_POLICY, 0));
        else out[i] = (int32_t)tag_preferred_mask(T, s, lo, la);
    }
}
// tag.py:303-316 with the heuristic: a = choice(_generate_preferred(history)); ob, rw, done = step(a); history.append(a, ob)
template <int NOPP>
__global__ void __launch_bounds__(POMDP_THREADS)
pomdp_tag_rollout_preferred_kernel(const __grid_constant__ TagDev p, const void* __restrict__ g_table, const int32_t* state,
                                   int32_t* __restrict__ last_obs, int32_t* __restrict__ last_action,
                                   const int32_t* __restrict__ first_action, int32_t* final_state, double* __restrict__ ret,
                                   int32_t* __restrict__ steps, int32_t* __restrict__ flags, int64_t n, uint64_t goff,
                                   const __grid_constant__ PhiloxKey seed, uint32_t ctr0, int32_t max_steps, double gamma) {
    typedef TagEnvT<NOPP> Env;
    extern __shared__ __align__(128) unsigned char smem_table[];
    __shared__ alignas(8) uint64_t bar;
    stage_table_sync<Env>(smem_table, g_table, NOPP == 1 ? (uint32_t)sizeof(TagTables) : TAG_TABLES_BASE_BYTES, &bar);   // one opponent: + the LUT
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nthreads) {
        uint32_t s = (uint32_t)state[i];
        int32_t lo = last_obs ? last_obs[i] : 0, la = last_action ? last_action[i] : -1;
        RolloutAcc acc;
        tag_rollout_preferred1<NOPP>(p, smem_table, s, seed, goff + (uint64_t)i, ctr0, max_steps, gamma, first_action != nullptr,
                                     first_action ? first_action[i] : 0, lo, la, acc);
        if (final_state) final_state[i] = (int32_t)s;
        ret[i] = acc.ret; steps[i] = acc.steps; flags[i] = acc.flags;
        if (last_obs) last_obs[i] = lo;
        if (last_action) last_action[i] = la;
    }
}

__global__ void __launch_bounds__(POMDP_THREADS)
pomdp_battleship_obs_prob_kernel(const __grid_constant__ ShipDev p, const int32_t* __restrict__ state,
                                 const int32_t* __restrict__ action, const int32_t* __restrict__ obs,
                                 double* __restrict__ prob, int64_t n) {
    const bool vec = (reinterpret_cast<uintptr_t>(state) & 15) == 0;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nthreads) {
        uint32_t w[SHIP_WORDS];
        ship_load(state, i, vec, w);
        prob[i] = battleship_obs_prob(p, w, action[i], obs[i]);
    }
}
// battleship.py:157-165: bit c of the ceil(n_tiles / 32)-word mask = cell c not visited
__global__ void __launch_bounds__(POMDP_THREADS)
pomdp_battleship_legal_mask_kernel(const __grid_constant__ ShipDev p, const int32_t* __restrict__ state,
                                   uint32_t* __restrict__ mask, int64_t n) {
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    const int words = (p.n_tiles + 31) >> 5;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nthreads) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int lim = p.n_tiles - 32 * k;
            const uint32_t valid = lim >= 32 ? 0xFFFFFFFFu : (lim <= 0 ? 0u : ((1u << lim) - 1u));
            if (k < words) mask[i * words + k] = ~(uint32_t)state[i * SHIP_WORDS + 4 + k] & valid;
        }
    }
}

__global__ void __launch_bounds__(POMDP_THREADS)
pomdp_battleship_policy_kernel(const __grid_constant__ ShipDev p, const int32_t* __restrict__ state,
                               int32_t* __restrict__ action, int64_t n, uint64_t goff,
                               const __grid_constant__ PhiloxKey seed, uint32_t step_ctr) {
    const bool vec = (reinterpret_cast<uintptr_t>(state) & 15) == 0;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nthreads) {
        uint32_t w[SHIP_WORDS];
        ship_load(state, i, vec, w);
        action[i] = battleship_policy(p, w, draw_word(seed, goff + (uint64_t)i, step_ctr, DOMAIN_POLICY, 0));
    }
}
__global__ void __launch_bounds__(POMDP_THREADS)
pomdp_battleship_rollout_kernel(const __grid_constant__ ShipDev p, const int32_t* state,
                                const int32_t* __restrict__ first_action, int32_t* final_state,
                                double* __restrict__ ret, int32_t* __restrict__ steps, int32_t* __restrict__ flags,
                                int64_t n, uint64_t goff, const __grid_constant__ PhiloxKey seed, uint32_t ctr0,
                                int32_t max_steps, double gamma) {
    const bool vec = ((reinterpret_cast<uintptr_t>(state) | reinterpret_cast<uintptr_t>(final_state)) & 15) == 0;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nthreads) {
        uint32_t w[SHIP_WORDS], w2[SHIP_WORDS];
        ship_load(state, i, vec, w);
        RolloutAcc acc;
        acc.init((w[3] >> 31) != 0);
        for (int32_t t = 0; t < max_steps && !(w[3] >> 31); ++t) {
            const int32_t a = (t == 0 && first_action)
                                  ? first_action[i]
                                  : battleship_policy(p, w, draw_word(seed, goff + (uint64_t)i, ctr0 + (uint32_t)t, DOMAIN_POLICY, 0));
            int32_t ob, fl;
            float rw;
            battleship_step(p, w, a, w2, ob, rw, fl);
#pragma unroll
            for (int k = 0; k < SHIP_WORDS; ++k) w[k] = w2[k];
            acc.add((double)rw, gamma, fl);
        }
        if (final_state) {
            if (vec) {
                uint4* dst = reinterpret_cast<uint4*>(final_state + i * SHIP_WORDS);
                __stcs(dst, make_uint4(w[0], w[1], w[2], w[3]));
                __stcs(dst + 1, make_uint4(w[4], w[5], w[6], w[7]));
            } else {
#pragma unroll
                for (int k = 0; k < SHIP_WORDS; ++k) final_state[i * SHIP_WORDS + k] = (int32_t)w[k];
            }
        }
        ret[i] = acc.ret; steps[i] = acc.steps; flags[i] = acc.flags;
    }
}

// --------------------------------------------------------------- BattleShip step ----
// One thread per board; the CTA's tile of boards (POMDP_THREADS x 32 B) travels through
// shared memory with TMA bulk copies in both directions.
__global__ void __launch_bounds__(POMDP_THREADS)
pomdp_battleship_step_kernel(const __grid_constant__ ShipDev p, const int32_t* state,
                             const int32_t* __restrict__ action, int32_t* next_state, int32_t* __restrict__ obs,
                             float* __restrict__ reward, int32_t* __restrict__ flags, int64_t n) {
    __shared__ alignas(128) uint32_t tile[POMDP_THREADS * SHIP_WORDS];
    __shared__ alignas(8) uint64_t bar;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    __syncthreads();
    const int64_t n_tiles = (n + POMDP_THREADS - 1) / POMDP_THREADS;
    uint32_t parity = 0;
    for (int64_t tix = blockIdx.x; tix < n_tiles; tix += gridDim.x) {
        const int64_t base = tix * POMDP_THREADS;
        const int cnt = (int)min((int64_t)POMDP_THREADS, n - base);
        const uint32_t bytes = (uint32_t)cnt * SHIP_WORDS * 4u;
        if (threadIdx.x == 0) {
            mbar_expect_tx(&bar, bytes);
            tma_bulk_g2s(tile, state + base * SHIP_WORDS, bytes, &bar);
        }
        const int64_t i = base + threadIdx.x;
        int32_t a = 0;
        if (threadIdx.x < cnt) a = __ldcs(action + i);
        mbar_wait(&bar, parity);
        parity ^= 1;
        if (threadIdx.x < cnt) {
            uint32_t w[SHIP_WORDS], w2[SHIP_WORDS];
            uint4* mine = reinterpret_cast<uint4*>(tile + threadIdx.x * SHIP_WORDS);
            const uint4 lo = mine[0], hi = mine[1];
            w[0] = lo.x; w[1] = lo.y; w[2] = lo.z; w[3] = lo.w; w[4] = hi.x; w[5] = hi.y; w[6] = hi.z; w[7] = hi.w;
            int32_t ob, fl; float rw;
            battleship_step(p, w, a, w2, ob, rw, fl);
            mine[0] = make_uint4(w2[0], w2[1], w2[2], w2[3]);
            mine[1] = make_uint4(w2[4], w2[5], w2[6], w2[7]);
            __stcs(obs + i, ob); __stcs(reward + i, rw); __stcs(flags + i, fl);
        }
        fence_proxy_async_smem();     // generic-proxy smem writes -> visible to the bulk-copy engine
        __syncthreads();
        if (threadIdx.x == 0) {
            tma_bulk_s2g(next_state + base * SHIP_WORDS, tile, bytes);
            tma_commit();
            tma_wait_read0();         // tile may be overwritten once the engine has read it
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) tma_wait_all0();
}

// Fallback for state pointers that are not 16-byte aligned (bulk copies need that).
__global__ void __launch_bounds__(POMDP_THREADS)
pomdp_battleship_step_plain_kernel(const __grid_constant__ ShipDev p, const int32_t* state,
                                   const int32_t* __restrict__ action, int32_t* next_state, int32_t* __restrict__ obs,
                                   float* __restrict__ reward, int32_t* __restrict__ flags, int64_t n) {
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nthreads) {
        uint32_t w[SHIP_WORDS], w2[SHIP_WORDS];
#pragma unroll
        for (int k = 0; k < SHIP_WORDS; ++k) w[k] = (uint32_t)state[i * SHIP_WORDS + k];
        int32_t ob, fl; float rw;
        battleship_step(p, w, action[i], w2, ob, rw, fl);
#pragma unroll
        for (int k = 0; k < SHIP_WORDS; ++k) next_state[i * SHIP_WORDS + k] = (int32_t)w2[k];
        obs[i] = ob; reward[i] = rw; flags[i] = fl;
    }
}

// --------------------------------------------------------------- BattleShip reset ---
// One warp per env (BASELINE.json: "warp-per-env ship scan").  For each ship the 32 lanes
// test the 4*n_tiles (pos, dir) candidates c = 4*pos + dir, lane l taking c = l + 32 j;
// ballots give, per j, the accepted set in increasing c; the k-th accepted candidate is
// taken with k = floor(u * count), u from draw slot = ship index.
__global__ void __launch_bounds__(POMDP_THREADS)
pomdp_battleship_reset_scan_kernel(const __grid_constant__ ShipDev p, int32_t* __restrict__ state,
                                   int32_t* __restrict__ obs, int32_t* __restrict__ flags,
                                   const uint8_t* __restrict__ mask, int64_t n, uint64_t goff,
                                   const __grid_constant__ PhiloxKey seed, uint32_t step_ctr) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int n_cand = 4 * p.n_tiles;
    const int n_iter = (n_cand + 31) >> 5;    // <= 15
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
        if (mask && !mask[i]) continue;       // warp-uniform
        ShipState st;
        st.occ = b128(0, 0); st.vis = b128(0, 0); st.remaining = 0; st.done = false;
        bool ok_all = true;
        int ship = 0;
        for (int length = p.max_len; length >= 2; --length, ++ship) {
            const B128 blocked = ship_blocked(p, st.occ);
            uint32_t mine = 0;                // bit j: candidate lane + 32 j is accepted
            int total = 0;
            for (int j = 0; j < n_iter; ++j) {
                const int c = lane + 32 * j;
                const bool ok = c < n_cand && ship_candidate_ok(p, blocked, c >> 2, c & 3, length);
                mine |= (uint32_t)ok << j;
                total += __popc(__ballot_sync(0xffffffffu, ok));
            }
            if (total == 0) { ok_all = false; break; }   // the reference would loop forever
            const uint32_t w = ShipDraw(seed, goff + (uint64_t)i, step_ctr)(ship);                            // warp-uniform
            int k = (int)rand_below(w, (uint32_t)total);
            int chosen = -1;
            for (int j = 0; j < n_iter; ++j) {
                const uint32_t m = __ballot_sync(0xffffffffu, (mine >> j) & 1u);
                const int cnt = __popc(m);
                if (chosen < 0) {
                    if (k < cnt) chosen = 32 * j + (int)__fns(m, 0, k + 1);
                    else k -= cnt;
                }
            }
            ship_mark(p, st, chosen >> 2, chosen & 3, length);
        }
        uint32_t w8[SHIP_WORDS];
        ship_pack(st, w8);
        if (lane < SHIP_WORDS) {
            uint32_t v = w8[0];
#pragma unroll
            for (int k = 1; k < SHIP_WORDS; ++k) if (lane == k) v = w8[k];
            state[i * SHIP_WORDS + lane] = (int32_t)v;
        }
        if (lane == 8 && obs) obs[i] = 0;                                  // battleship.py:137
        if (lane == 9 && flags) flags[i] = ok_all ? 0 : FLAG_BAD_STATE;
    }
}

// One THREAD per env: the 4 * n_tiles candidates of a ship are four 128-bit masks built with ~length shifts each
// (pomdp_core.h: ship_valid_starts), so a placement costs ~1e3 instructions instead of the warp scan's ~5e4.
__global__ void __launch_bounds__(POMDP_THREADS)
pomdp_battleship_reset_bitboard_kernel(const __grid_constant__ ShipDev p, int32_t* __restrict__ state,
                                       int32_t* __restrict__ obs, int32_t* __restrict__ flags,
                                       const uint8_t* __restrict__ mask, int64_t n, uint64_t goff,
                                       const __grid_constant__ PhiloxKey seed, uint32_t step_ctr) {
    const bool vec = (reinterpret_cast<uintptr_t>(state) & 15) == 0;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nthreads) {
        if (mask && !mask[i]) continue;
        ShipState st;
        const bool ok = battleship_reset_bitboard(p, seed, goff + (uint64_t)i, step_ctr, st);
        uint32_t w8[SHIP_WORDS];
        ship_pack(st, w8);
        if (vec) {
            uint4* dst = reinterpret_cast<uint4*>(state + i * SHIP_WORDS);
            __stcs(dst, make_uint4(w8[0], w8[1], w8[2], w8[3]));
            __stcs(dst + 1, make_uint4(w8[4], w8[5], w8[6], w8[7]));
        } else {
#pragma unroll
            for (int k = 0; k < SHIP_WORDS; ++k) state[i * SHIP_WORDS + k] = (int32_t)w8[k];
        }
        if (obs) obs[i] = 0;                                               // battleship.py:137
        if (flags) flags[i] = ok ? 0 : FLAG_BAD_STATE;
    }
}

// One thread per env, placements of the first two ships read from the host-built tables (pomdp_core.h:
// battleship_reset_table) -- one Philox call and two dependent table reads per board instead of the bitboard scan.
// kBulk (state 16-byte aligned, no mask): the CTA's tile of boards is assembled in shared memory and leaves with ONE
// TMA bulk store per tile, like the step kernel's; otherwise every thread stores its own 32 bytes.
// kLean: the stock two-ship game -- no scan code in the kernel.  CTAs are 128 threads (4 KB tiles): 2^18 boards are 2048
// tiles, which all fit on the machine at once, so every CTA runs its one chain of Philox -> record -> list -> store once.
#ifndef POMDP_SHIP_RESET_THREADS
#define POMDP_SHIP_RESET_THREADS 128
#endif
template <bool kBulk, bool kLean>
__global__ void __launch_bounds__(POMDP_SHIP_RESET_THREADS)
pomdp_battleship_reset_table_kernel(const __grid_constant__ ShipDev p, const unsigned char* __restrict__ tbl,
                                    int32_t* __restrict__ state, int32_t* __restrict__ obs, int32_t* __restrict__ flags,
                                    const uint8_t* __restrict__ mask, int64_t n, uint64_t goff,
                                    const __grid_constant__ PhiloxKey seed, uint32_t step_ctr) {
    constexpr int T = POMDP_SHIP_RESET_THREADS;
    __shared__ alignas(128) uint32_t tile[kBulk ? T * SHIP_WORDS : 4];
    const int64_t n_tiles = (n + T - 1) / T;
    for (int64_t tix = blockIdx.x; tix < n_tiles; tix += gridDim.x) {
        const int64_t base = tix * T;
        const int cnt = (int)min((int64_t)T, n - base);
        const int64_t i = base + threadIdx.x;
        if ((int)threadIdx.x < cnt && (kBulk || !mask || mask[i])) {
            ShipState st;
            const bool ok = battleship_reset_table<kLean>(p, tbl, ShipDraw(seed, goff + (uint64_t)i, step_ctr), st);
            uint32_t w8[SHIP_WORDS];
            ship_pack(st, w8);
            if (kBulk) {
                uint4* mine = reinterpret_cast<uint4*>(tile + threadIdx.x * SHIP_WORDS);
                mine[0] = make_uint4(w8[0], w8[1], w8[2], w8[3]);
                mine[1] = make_uint4(w8[4], w8[5], w8[6], w8[7]);
            } else if ((reinterpret_cast<uintptr_t>(state) & 15) == 0) {
                uint4* dst = reinterpret_cast<uint4*>(state + i * SHIP_WORDS);
                __stcs(dst, make_uint4(w8[0], w8[1], w8[2], w8[3]));
                __stcs(dst + 1, make_uint4(w8[4], w8[5], w8[6], w8[7]));
            } else {
#pragma unroll
                for (int k = 0; k < SHIP_WORDS; ++k) state[i * SHIP_WORDS + k] = (int32_t)w8[k];
            }
            if (obs) __stcs(obs + i, 0);                                   // battleship.py:137
            if (flags) __stcs(flags + i, ok ? 0 : (int32_t)FLAG_BAD_STATE);
        }
        if (kBulk) {
            fence_proxy_async_smem();     // generic-proxy smem writes -> visible to the bulk-copy engine
            __syncthreads();
            if (threadIdx.x == 0) {
                tma_bulk_s2g(state + base * SHIP_WORDS, tile, (uint32_t)cnt * SHIP_WORDS * 4u);
                tma_commit();
                if (tix + gridDim.x < n_tiles) tma_wait_read0();   // another tile follows: the engine must have read this one
            }
            if (tix + gridDim.x < n_tiles) __syncthreads();
        }
    }
    if (kBulk && threadIdx.x == 0) tma_wait_all0();
}

__global__ void __launch_bounds__(POMDP_THREADS)
pomdp_battleship_reset_rejection_kernel(const __grid_constant__ ShipDev p, int32_t* __restrict__ state,
                                        int32_t* __restrict__ obs, int32_t* __restrict__ flags,
                                        const uint8_t* __restrict__ mask, int64_t n, uint64_t goff,
                                        const __grid_constant__ PhiloxKey seed, uint32_t step_ctr) {
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nthreads) {
        if (mask && !mask[i]) continue;
        ShipState st;
        const bool ok = battleship_reset_rejection(p, seed, goff + (uint64_t)i, step_ctr, st, 4096);
        uint32_t w8[SHIP_WORDS];
        ship_pack(st, w8);
        uint4* dst = reinterpret_cast<uint4*>(state + i * SHIP_WORDS);
        if ((reinterpret_cast<uintptr_t>(state) & 15) == 0) {
            dst[0] = make_uint4(w8[0], w8[1], w8[2], w8[3]);
            dst[1] = make_uint4(w8[4], w8[5], w8[6], w8[7]);
        } else {
#pragma unroll
            for (int k = 0; k < SHIP_WORDS; ++k) state[i * SHIP_WORDS + k] = (int32_t)w8[k];
        }
        if (obs) obs[i] = 0;
        if (flags) flags[i] = ok ? 0 : FLAG_BAD_STATE;
    }
}

// ------------------------------------------------------------------- coord helpers ---
__global__ void __launch_bounds__(POMDP_THREADS)
pomdp_coord_kernel(int op, int xs, int ys, const int32_t* __restrict__ a, const int32_t* __restrict__ b,
                   int32_t* __restrict__ out, int64_t n) {
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nthreads) {
        switch (op) {
            case POMDP_COORD_GET_INDEX: out[i] = grid_get_index(xs, a[2 * i], a[2 * i + 1]); break;
            case POMDP_COORD_GET_COORD: out[2 * i] = a[i] % xs; out[2 * i + 1] = a[i] / xs; break;
            case POMDP_COORD_IS_INSIDE: out[i] = grid_is_inside(xs, ys, a[2 * i], a[2 * i + 1]); break;
            case POMDP_COORD_ADD_MOVE: {
                const int m = b[i];
                out[2 * i] = a[2 * i] + move_dx(m);
                out[2 * i + 1] = a[2 * i + 1] + move_dy(m);
                break;
            }
            case POMDP_COORD_L1: out[i] = l1_distance(a[2 * i], a[2 * i + 1], b[2 * i], b[2 * i + 1]); break;
            case POMDP_COORD_TAG_GET_INDEX: {
                const int x = a[2 * i], y = a[2 * i + 1];
                out[i] = tag_is_inside(x, y) ? tag_get_index(x, y) : -1;
                break;
            }
            case POMDP_COORD_TAG_GET_COORD: {
                int x = -1, y = -1;
                if ((uint32_t)a[i] < (uint32_t)TAG_CELLS) tag_get_coord((uint32_t)a[i], x, y);
                out[2 * i] = x; out[2 * i + 1] = y;
                break;
            }
            case POMDP_COORD_TAG_IS_INSIDE: out[i] = tag_is_inside(a[2 * i], a[2 * i + 1]); break;
        }
    }
}

// ---------------------------------------------------------------- belief histogram ---
// Shared-memory histogram per CTA (<= 512 bins), then one 64-bit global atomic per non-empty bin per CTA; one
// 1024-thread CTA per SM, because that final step costs one same-address global atomic per bin per CTA
// (pomdp_core.h: belief_bins is the definition the tests check against).  Two kinds of bins:
//  * bit bins (Rock "rock i still good", Tiger door, BattleShip occupied cells, Network "machine m up"), which every
//    particle hits with probability ~1/2 -- as shared-memory atomics these are 32-way same-address conflicts, and as
//    warp ballots they are VOTE-throughput-bound (both measured: 30 us and 25-37 us for 2^22 Rock states).  Instead
//    every thread keeps PACKED BYTE COUNTERS in registers: four bits of the state are spread to the four bytes of
//    a word with one multiply ((x & 0xF) * 0x00204081 & 0x01010101) and added -- three instructions per four bins
//    per particle, no cross-lane traffic.  Counters are flushed (hardware warp reduction of the 16-bit halves, then
//    one shared-memory atomic per bin per warp) before a byte can overflow, normally once at the end;
//  * categorical bins (Rock agent cell, Tag agent/opponent cell): plain shared-memory atomics (many addresses).
// W = 1 and W = 2 states are read with 16-byte loads (four / two envs per thread per trip).
#ifndef POMDP_HIST_INFLIGHT
#define POMDP_HIST_INFLIGHT 4
#endif
#ifndef POMDP_HIST_PREFETCH
#define POMDP_HIST_PREFETCH 0
#endif
template <int KIND> struct HistShape;                        // NW = packed counter words per thread (4 bins each)
template <> struct HistShape<POMDP_KIND_ROCK> { static constexpr int NW = 4; };
template <> struct HistShape<POMDP_KIND_TAG> { static constexpr int NW = 1; };
template <> struct HistShape<POMDP_KIND_TIGER> { static constexpr int NW = 1; };
template <> struct HistShape<POMDP_KIND_NETWORK> { static constexpr int NW = 8; };
template <> struct HistShape<POMDP_KIND_BATTLESHIP> { static constexpr int NW = 30; };

__device__ __forceinline__ uint32_t spread4(uint32_t nibble) { return (nibble * 0x00204081u) & 0x01010101u; }        // bit i -> byte i
__device__ __forceinline__ uint32_t spread4_even(uint32_t byte) { return ((byte & 0x55u) * 0x00041041u) & 0x01010101u; }  // bit 2i -> byte i

// (Skipping the counter words past the bins in use -- Network-v0 has 10 machines, not 32 -- was measured and is SLOWER:
// the uniform branches cost the unrolled loop its instruction-level parallelism, profiles/r04g_hist_guard_variant.log.)
template <int KIND>
__device__ __forceinline__ void hist_one(int p0, bool valid, const uint32_t s[4], uint32_t* sh, uint32_t (&acc)[HistShape<KIND>::NW]) {
    if (KIND == POMDP_KIND_ROCK) {                           // bit 2i of `good` = rock i's status is +1 (code 01)
        const uint32_t good0 = (s[0] >> 8) & ~(s[0] >> 9) & 0x00555555u;   // rocks 0..11: bits 8..31 of word 0
        const uint32_t good1 = s[1] & ~(s[1] >> 1) & 0x00000055u;          // rocks 12..15: bits 0..7 of word 1
        acc[0] += spread4_even(good0);
        acc[1] += spread4_even(good0 >> 8);
        acc[2] += spread4_even(good0 >> 16);
        acc[3] += spread4_even(good1);
        if (valid) atomicAdd(&sh[p0 + (int)(s[0] & 0xFFu)], 1u);
    } else if (KIND == POMDP_KIND_TAG) {
        if (valid) {
            atomicAdd(&sh[s[0] & 31u], 1u);
            atomicAdd(&sh[TAG_CELLS + ((s[0] >> 5) & 31u)], 1u);
        }
    } else if (KIND == POMDP_KIND_TIGER) {
        acc[0] += valid ? (1u << (8u * (s[0] & 1u))) : 0u;
    } else if (KIND == POMDP_KIND_NETWORK) {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += spread4((s[0] >> (4 * j)) & 0xFu);
    } else {                                                 // BattleShip: 120 occupied bits over four words
#pragma unroll
        for (int j = 0; j < 30; ++j) acc[j] += spread4((s[j >> 3] >> (4 * (j & 7))) & 0xFu);
    }
}
// adds the warp's packed byte counters to the shared histogram (bins 4j + c < n_bits) and clears them: the bytes of
// word j are widened to two words of 16-bit fields (32 lanes x 255 fits), each summed over the warp by ONE hardware
// reduction (REDUX via __reduce_add_sync; a plain 32-bit add never carries between the fields), and lane j keeps
// word j's totals, so the shared-memory atomics of all words are issued together, four per lane.
template <int NW>
__device__ __forceinline__ void hist_flush(uint32_t (&acc)[NW], uint32_t* sh, int n_bits, int lane, uint32_t weight = 1u) {
    static_assert(NW <= 32, "one lane per counter word");
    uint32_t my_lo = 0, my_hi = 0;
#pragma unroll
    for (int j = 0; j < NW; ++j) {
        if (4 * j >= n_bits) break;
        const uint32_t lo = __reduce_add_sync(0xffffffffu, acc[j] & 0x00FF00FFu);          // bytes 0, 2
        const uint32_t hi = __reduce_add_sync(0xffffffffu, (acc[j] >> 8) & 0x00FF00FFu);   // bytes 1, 3
        if (lane == j) { my_lo = lo; my_hi = hi; }
        acc[j] = 0;
    }
    if (4 * lane < n_bits) {
        const uint32_t c[4] = {my_lo & 0xFFFFu, my_hi & 0xFFFFu, my_lo >> 16, my_hi >> 16};
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (c[k] && 4 * lane + k < n_bits) atomicAdd(&sh[4 * lane + k], c[k] * weight);
    }
}
// ---- bit bins by carry-save addition (Harley-Seal) ------------------------------------------------------------------
// Rock's "rock i still good" and Network's "machine m up" are one bit per bin per particle.  Instead of spreading every
// particle's bits into byte counters (5 instructions per four bins per particle), the words of a trip are summed
// VERTICALLY: a carry-save adder (two LOP3) turns three words of weight w into one of weight w and one of weight 2w, and
// a tree of 15 of them reduces the 16 words a thread loads per trip (7 for the 8 two-word states) to running planes of
// weight 1, 2, 4, 8 plus ONE carry-out word of weight 16 (8) -- only that word is spread into the byte counters, whose
// unit becomes 16 (8) particles.  The running planes are added once, at the end.
__device__ __forceinline__ void csa(uint32_t& hi, uint32_t& lo, uint32_t a, uint32_t b, uint32_t c) {
    const uint32_t u = a ^ b;
    hi = (a & b) | (u & c);
    lo = u ^ c;
}
template <int KIND>
__device__ __forceinline__ uint32_t hist_bitword(uint32_t s0, uint32_t s1) {
    if (KIND == POMDP_KIND_ROCK)    // bit 2i = rock i's status is +1 (code 01); rocks 12..15 (word 1) in bits 24..30
        return ((s0 >> 8) & ~(s0 >> 9) & 0x00555555u) | ((s1 & ~(s1 >> 1) & 0x00000055u) << 24);
    return s0;                      // Network: bit m = machine m is up
}
template <int KIND>
__device__ __forceinline__ void hist_add_plane(uint32_t x, uint32_t (&acc)[HistShape<KIND>::NW]) {
    if (KIND == POMDP_KIND_ROCK) {
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] += spread4_even(x >> (8 * j));
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += spread4((x >> (4 * j)) & 0xFu);
    }
}

// The end of every kernel that counts a belief histogram (pomdp_belief_hist_kernel, pomdp_step_hist_kernel): the CTA's
// counts are complete in shared memory (`sh`, after a __syncthreads); called by ALL threads of the CTA.
// Fused all-reduce (pomdp_belief_hist_allreduce): `hist` is this rank's scratch -- hist[bins] a ticket counter,
// hist[bins + 1] the number of calls made so far.  The CTA that takes the last ticket owns the rank's complete counts.
// Call e (1, 2, ...) uses result slot (e - 1) & 1 of every rank's symmetric buffer [slot 0 | slot 1 | arrivals]:
//   1. clear this rank's OTHER slot for call e + 1 (nobody adds into it before having seen this rank's arrival of
//      call e, and its previous contents -- the result of call e - 1 -- were handed out by that call),
//   2. add the counts into EVERY rank's slot through the peer mappings: system-scope reductions over NVLink/NVSwitch,
//   3. announce the arrival in row [rank] of every peer's arrival counters (release, system scope: the reductions are
//      ordered before it) and wait until all counters of this rank's own row block have reached e (acquire),
//   4. copy the slot -- now the GLOBAL counts -- to hist_out.
// One kernel instead of zero-fill + histogram + a collective; the epoch lives in device memory, so the launch is
// identical call after call and can be replayed from a CUDA graph.
__device__ __forceinline__ void hist_finish(const uint32_t* __restrict__ sh, unsigned long long* __restrict__ hist, int bins,
                                            unsigned long long* const* __restrict__ peers, int world, int rank, int wait,
                                            unsigned long long* __restrict__ hist_out) {
    // Self-cleaning local call (pomdp_belief_hist_once, a local sink of pomdp_E_step_hist: no peers): ONE atomic per bin per
    // CTA both adds the CTA's count (low 48 bits) and counts the CTA's arrival at that bin (high 16 bits); the thread whose
    // atomic returns the last arrival owns the bin's complete count: it writes hist_out[b] and clears the scratch word.  One
    // L2 round trip ends the kernel -- no fence, ticket and read-back in sequence (three round trips, measured ~1.4 us) --
    // no zero-fill launch precedes it, and the scratch is all zero again after it.
    if (!peers && hist_out) {
        const unsigned long long arrivals = (unsigned long long)gridDim.x - 1ull;
        for (int b = threadIdx.x; b < bins; b += blockDim.x) {
            const unsigned long long old = atomicAdd(&hist[b], (1ull << 48) | (unsigned long long)sh[b]);
            if ((old >> 48) == arrivals) {
                hist_out[b] = (old + sh[b]) & ((1ull << 48) - 1ull);
                hist[b] = 0ull;
            }
        }
        return;
    }
    for (int b = threadIdx.x; b < bins; b += blockDim.x)
        if (sh[b]) atomicAdd(&hist[b], (unsigned long long)sh[b]);
    if (peers) {
        __shared__ bool last;
        __shared__ unsigned long long epoch_s;
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
            last = atomicAdd(&hist[bins], 1ull) == (unsigned long long)gridDim.x - 1ull;
            epoch_s = hist[bins + 1] + 1ull;
        }
        __syncthreads();
        if (last) {
            __threadfence();
            const unsigned long long epoch = epoch_s;
            const int64_t slot_off = (int64_t)((epoch - 1ull) & 1ull) * POMDP_HIST_MAX_BINS * 8;
            const int64_t other_off = POMDP_HIST_MAX_BINS * 8 - slot_off;
            const int64_t signal_off = 2 * POMDP_HIST_MAX_BINS * 8;
            char* own = reinterpret_cast<char*>(peers[rank]);
            for (int b = threadIdx.x; b < POMDP_HIST_MAX_BINS; b += blockDim.x)
                reinterpret_cast<unsigned long long*>(own + other_off)[b] = 0ull;
            for (int b = threadIdx.x; b < bins; b += blockDim.x) {
                const unsigned long long v = atomicExch(&hist[b], 0ull);
                if (v)
                    for (int r = 0; r < world; ++r)
                        atomicAdd_system(reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(peers[r]) + slot_off) + b, v);
            }
            if (wait) {
                __threadfence_system();
                __syncthreads();
                if ((int)threadIdx.x < world) {
                    const int r = (int)threadIdx.x;
                    unsigned long long* there = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(peers[r]) + signal_off) + rank;
                    const unsigned long long* here = reinterpret_cast<const unsigned long long*>(own + signal_off) + r;
                    asm volatile("red.release.sys.global.add.u64 [%0], %1;" ::"l"(there), "l"(1ull) : "memory");
                    unsigned long long seen = 0, t0, t1;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
                    for (;;) {
                        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(here) : "memory");
                        if (seen >= epoch) break;
                        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                        if (t1 - t0 > 10000000000ull) __trap();            // a peer never made the call: fail, do not hang
                    }
                }
                __syncthreads();
            }
            if (hist_out)
                for (int b = threadIdx.x; b < bins; b += blockDim.x)
                    hist_out[b] = __ldcv(reinterpret_cast<const unsigned long long*>(own + slot_off) + b);
            __syncthreads();
            if (threadIdx.x == 0) { hist[bins] = 0ull; hist[bins + 1] = epoch; }
        }
    }
}

template <int KIND, bool CSA>
__global__ void __launch_bounds__(1024, 1)      // one CTA per SM is all the host launches: no reason to squeeze into 32 registers
pomdp_belief_hist_kernel(int p0, int p1, const int32_t* __restrict__ state, int words, int64_t n,
                         unsigned long long* __restrict__ hist, int bins,
                         unsigned long long* const* __restrict__ peers, int world, int rank, int wait,
                         unsigned long long* __restrict__ hist_out) {
    constexpr int NW = HistShape<KIND>::NW;
    __shared__ uint32_t sh[POMDP_HIST_MAX_BINS];
    for (int b = threadIdx.x; b < bins; b += blockDim.x) sh[b] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int n_bits = KIND == POMDP_KIND_TAG ? 0 : KIND == POMDP_KIND_TIGER ? 2 : p0;
    uint32_t acc[NW];
#pragma unroll
    for (int j = 0; j < NW; ++j) acc[j] = 0;
    int pending = 0;                                         // particles added since the last flush (warp-uniform)
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    int64_t scalar_from = 0;
    if ((words == 1 || words == 2) && (reinterpret_cast<uintptr_t>(state) & 15) == 0) {
        // one 16-byte load per thread per trip: four W = 1 states or two W = 2 states
        const int lg = words == 1 ? 2 : 1;
        const int64_t n_groups = n >> lg;
        const int64_t g_round = (n_groups + 31) & ~(int64_t)31;          // whole warps iterate together (flush shuffles)
        // kInFlight independent 16-byte loads per thread per trip: with one, a 1024-thread CTA per SM keeps only 16 KB in
        // flight and the kernel waits on DRAM latency (long-scoreboard stalls, 20-30 % of the HBM peak on a pure read)
        constexpr int kInFlight = POMDP_HIST_INFLIGHT;
        // kPrefetch: the loads of a thread's NEXT trip are issued before the current trip is counted (software pipelining,
        // as in the step kernels).  Measured with 2, 4 and 8 loads per trip, with and without (profiles/r05e_hist_loop_variants.log):
        // nothing moves below 2^22 states -- there the kernel is a fixed cost (launch and the ending: ~4 us with the fence +
        // ticket + read-back ending of that measurement, ~2.6 us with hist_finish's one-atomic ending) plus 0.7 us per 2^20
        // one-word states, i.e. the marginal rate is already the HBM rate -- and at 2^25 four plain loads per trip are the
        // fastest (49 us; 54-64 us for the others).
        constexpr bool kPrefetch = POMDP_HIST_PREFETCH != 0;
        // CSA: chosen by the host only when a thread makes many trips (measured: 2^25 two-word states 57.0 -> 48.6 us, but
        // slower at 2^22 and below, where a thread makes two trips and the final planes cost more than they save)
        constexpr bool kCsa = CSA && kInFlight == 4 && (KIND == POMDP_KIND_ROCK || KIND == POMDP_KIND_NETWORK);   // the tree below sums 16 words
        uint32_t ones = 0, twos = 0, fours = 0, eights = 0;              // running planes of the vertical sum (kCsa)
        int4 v[kInFlight], nv[kInFlight];
        bool valid[kInFlight], nvalid[kInFlight];
        auto load_trip = [&](int64_t g0, int4 (&vv)[kInFlight], bool (&ok)[kInFlight]) {
#pragma unroll
            for (int u = 0; u < kInFlight; ++u) {
                const int64_t g = g0 + u * nthreads;
                ok[u] = g < n_groups;
                vv[u] = make_int4(0, 0, 0, 0);
                if (ok[u]) vv[u] = ld_stream4(state + (g << 2));
            }
        };
        if (kPrefetch && tid < g_round) load_trip(tid, v, valid);
        for (int64_t g0 = tid; g0 < g_round; g0 += kInFlight * nthreads) {
            if (kPrefetch) {
                if (g0 + kInFlight * nthreads < g_round) load_trip(g0 + kInFlight * nthreads, nv, nvalid);
            } else {
                load_trip(g0, v, valid);
            }
            if (kCsa) {
                uint32_t w[4 * kInFlight];                               // an invalid group contributes zero words
#pragma unroll
                for (int u = 0; u < kInFlight; ++u) {
                    const uint32_t e[4] = {(uint32_t)v[u].x, (uint32_t)v[u].y, (uint32_t)v[u].z, (uint32_t)v[u].w};
                    if (words == 1) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            w[4 * u + j] = hist_bitword<KIND>(e[j], 0u);
                            if (KIND == POMDP_KIND_ROCK && valid[u]) atomicAdd(&sh[p0 + (int)(e[j] & 0xFFu)], 1u);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            w[2 * u + j] = hist_bitword<KIND>(e[2 * j], e[2 * j + 1]);
                            if (KIND == POMDP_KIND_ROCK && valid[u]) atomicAdd(&sh[p0 + (int)(e[2 * j] & 0xFFu)], 1u);
                        }
                    }
                }
                uint32_t twosA, twosB, foursA, foursB, eightsA;
                csa(twosA, ones, ones, w[0], w[1]);
                csa(twosB, ones, ones, w[2], w[3]);
                csa(foursA, twos, twos, twosA, twosB);
                csa(twosA, ones, ones, w[4], w[5]);
                csa(twosB, ones, ones, w[6], w[7]);
                csa(foursB, twos, twos, twosA, twosB);
                csa(eightsA, fours, fours, foursA, foursB);
                if (words == 1) {
                    uint32_t eightsB, sixteens;
                    csa(twosA, ones, ones, w[8], w[9]);
                    csa(twosB, ones, ones, w[10], w[11]);
                    csa(foursA, twos, twos, twosA, twosB);
                    csa(twosA, ones, ones, w[12], w[13]);
                    csa(twosB, ones, ones, w[14], w[15]);
                    csa(foursB, twos, twos, twosA, twosB);
                    csa(eightsB, fours, fours, foursA, foursB);
                    csa(sixteens, eights, eights, eightsA, eightsB);
                    hist_add_plane<KIND>(sixteens, acc);                 // unit of the byte counters: 16 particles
                } else {
                    hist_add_plane<KIND>(eightsA, acc);                  // eight two-word states per trip: unit 8
                }
                if (++pending > 254) { hist_flush<NW>(acc, sh, n_bits, lane, words == 1 ? 16u : 8u); pending = 0; }
            } else {
#pragma unroll
                for (int u = 0; u < kInFlight; ++u) {
                    if (g0 + u * nthreads >= g_round) break;             // warp-uniform: g_round is a multiple of 32
                    const uint32_t e[4] = {(uint32_t)v[u].x, (uint32_t)v[u].y, (uint32_t)v[u].z, (uint32_t)v[u].w};
                    if (words == 1) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint32_t s[4] = {e[j], 0u, 0u, 0u};
                            hist_one<KIND>(p0, valid[u], s, sh, acc);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const uint32_t s[4] = {e[2 * j], e[2 * j + 1], 0u, 0u};
                            hist_one<KIND>(p0, valid[u], s, sh, acc);
                        }
                    }
                    pending += 4;
                    if (pending > 251) { hist_flush<NW>(acc, sh, n_bits, lane); pending = 0; }
                }
            }
            if (kPrefetch) {
#pragma unroll
                for (int u = 0; u < kInFlight; ++u) { v[u] = nv[u]; valid[u] = nvalid[u]; }
            }
        }
        if (kCsa) {                                  // the carry-outs, then the running planes in ONE flush (a byte holds <= 15)
            hist_flush<NW>(acc, sh, n_bits, lane, words == 1 ? 16u : 8u);
            uint32_t t[NW];
#pragma unroll
            for (int j = 0; j < NW; ++j) t[j] = 0;
            hist_add_plane<KIND>(ones, acc);
            hist_add_plane<KIND>(twos, t);
#pragma unroll
            for (int j = 0; j < NW; ++j) { acc[j] += 2u * t[j]; t[j] = 0; }
            hist_add_plane<KIND>(fours, t);
#pragma unroll
            for (int j = 0; j < NW; ++j) { acc[j] += 4u * t[j]; t[j] = 0; }
            hist_add_plane<KIND>(eights, t);         // stays zero for two-word states
#pragma unroll
            for (int j = 0; j < NW; ++j) acc[j] += 8u * t[j];
            hist_flush<NW>(acc, sh, n_bits, lane, 1u);
            pending = 0;
        }
        scalar_from = n_groups << lg;
    }
    const int64_t rem = n - scalar_from;
    const int64_t r_round = (rem + 31) & ~(int64_t)31;
    const bool aligned8 = (reinterpret_cast<uintptr_t>(state) & 7) == 0;
    const bool aligned16 = (reinterpret_cast<uintptr_t>(state) & 15) == 0;
    for (int64_t r = tid; r < r_round; r += nthreads) {
        const int64_t i = scalar_from + r;
        const bool valid = r < rem;
        uint32_t s[4] = {0u, 0u, 0u, 0u};
        if (valid) {
            if (words == 1) s[0] = (uint32_t)__ldcs(state + i);
            else if (words == 2 && aligned8) { const int2 v = __ldcs(reinterpret_cast<const int2*>(state) + i); s[0] = (uint32_t)v.x; s[1] = (uint32_t)v.y; }
            else if (words == SHIP_WORDS && aligned16) {     // BattleShip: the four occupied words of the 32-byte board
                const int4 v = __ldcs(reinterpret_cast<const int4*>(state) + 2 * i);
                s[0] = (uint32_t)v.x; s[1] = (uint32_t)v.y; s[2] = (uint32_t)v.z; s[3] = (uint32_t)v.w & 0x00FFFFFFu;
            } else {
                for (int k = 0; k < 4 && k < words; ++k) s[k] = (uint32_t)state[i * words + k];
                if (words == SHIP_WORDS) s[3] &= 0x00FFFFFFu;
            }
        }
        hist_one<KIND>(p0, valid, s, sh, acc);
        if (++pending > 254) { hist_flush<NW>(acc, sh, n_bits, lane); pending = 0; }
    }
    hist_flush<NW>(acc, sh, n_bits, lane);
    __syncthreads();
    (void)p1;
    hist_finish(sh, hist, bins, peers, world, rank, wait, hist_out);
}


// ------------------------------------------------- step with the belief histogram in its epilogue ---
// the counts of the NEXT states are accumulated by the step kernel itself -- the states are in registers
// when they are counted, so the particle set is never read back from HBM by a second kernel -- and the all-reduce of the
// counts can ride in the same launch (hist_finish: local move to hist_out, or the additions into every rank's buffer over
// NVLink peer memory).  Same loads, stores, draws and transition as pomdp_step_kernel; the loop runs whole warps
// together (an out-of-range lane idles) because the flush of the packed byte counters is a warp-wide reduction.
__device__ __forceinline__ void state_hist_words(uint32_t s, uint32_t w[4]) { w[0] = s; w[1] = 0u; w[2] = 0u; w[3] = 0u; }
__device__ __forceinline__ void state_hist_words(uint64_t s, uint32_t w[4]) { w[0] = (uint32_t)s; w[1] = (uint32_t)(s >> 32); w[2] = 0u; w[3] = 0u; }

template <class Env, bool kVec>
__global__ void __launch_bounds__(POMDP_STEP_THREADS, POMDP_STEP_MINB)
pomdp_step_hist_kernel(const __grid_constant__ typename Env::Params p, const void* __restrict__ g_table,
                       const int32_t* state, const int32_t* __restrict__ action, int32_t* next_state,
                       int32_t* __restrict__ obs, float* __restrict__ reward, int32_t* __restrict__ flags, int64_t n,
                       uint64_t goff, const __grid_constant__ PhiloxKey seed, uint32_t step_ctr, uint32_t table_bytes,
                       unsigned long long* __restrict__ hist, int bins, unsigned long long* const* __restrict__ peers,
                       int world, int rank, int wait, unsigned long long* __restrict__ hist_out) {
    typedef typename Env::State S;
    constexpr int KIND = Env::kHistKind;
    constexpr int NW = HistShape<KIND>::NW;
    extern __shared__ __align__(128) unsigned char smem_table[];
    __shared__ alignas(8) uint64_t bar;
    __shared__ uint32_t sh[POMDP_HIST_MAX_BINS];
    const unsigned char* lut = smem_table;
    for (int b = threadIdx.x; b < bins; b += blockDim.x) sh[b] = 0;
    if (Env::kTable && threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
        mbar_expect_tx(&bar, table_bytes);
        tma_bulk_g2s(smem_table, g_table, table_bytes, &bar);
    }
    __syncthreads();           // barrier object initialised and counters cleared before anyone touches them
    stage_param_table<Env>(smem_table, p);

    pdl_wait();
    pdl_launch_dependents();
    const int hp0 = Env::hist_p0(p);
    const int n_bits = KIND == POMDP_KIND_TAG ? 0 : KIND == POMDP_KIND_TIGER ? 2 : hp0;
    const int lane = threadIdx.x & 31;
    uint32_t acc[NW];
#pragma unroll
    for (int j = 0; j < NW; ++j) acc[j] = 0;
    int pending = 0;                                         // particles counted since the last flush (warp-uniform)
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    bool table_ready = !Env::kTable;
    if (kVec) {
        const int64_t n_groups = n >> 2;
        const int64_t g_round = (n_groups + 31) & ~(int64_t)31;
        const uint64_t group0 = goff >> 2;
        int64_t g = tid;
        StateVec<S> cur_s;
        cur_s.zero();
        int4 cur_a = make_int4(0, 0, 0, 0);
        if (g < n_groups) {
            cur_s.load(state, g << 2);
            cur_a = ld_stream4(action + (g << 2));
        }
        if (!table_ready) { mbar_wait(&bar, 0); table_ready = true; }
        while (g < g_round) {
            const int64_t gn = g + nthreads;
            StateVec<S> nxt_s = cur_s;
            int4 nxt_a = cur_a;
            if (gn < n_groups) {
                nxt_s.load(state, gn << 2);
                nxt_a = ld_stream4(action + (gn << 2));
            }
            if (g < n_groups) {
                const int64_t i = g << 2;
                S s[4], s2[4];
                cur_s.unpack(s);
                const int32_t a[4] = {cur_a.x, cur_a.y, cur_a.z, cur_a.w};
                int32_t ob[4], fl[4];
                float rw[4];
                Env::step4(p, lut, s, a, seed, group0 + (uint64_t)g, step_ctr, s2, ob, rw, fl);
                StateVec<S>::store(next_state, i, s2);
                st_stream4(obs + i, make_int4(ob[0], ob[1], ob[2], ob[3]));
                st_stream4(reward + i, make_float4(rw[0], rw[1], rw[2], rw[3]));
                st_stream4(flags + i, make_int4(fl[0], fl[1], fl[2], fl[3]));
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint32_t w[4];
                    state_hist_words(s2[j], w);
                    hist_one<KIND>(hp0, true, w, sh, acc);
                }
            }
            pending += 4;
            if (pending > 251) { hist_flush<NW>(acc, sh, n_bits, lane); pending = 0; }
            cur_s = nxt_s;
            cur_a = nxt_a;
            g = gn;
        }
        const int64_t i = (n_groups << 2) + tid;             // tail (n % 4 envs): the first few threads of the grid
        if (i < n) {
            S s2; int32_t ob, fl; float rw;
            Env::step1(p, lut, load_state1(state, i, S()), action[i], seed, goff + (uint64_t)i, step_ctr, s2, ob, rw, fl);
            store_state1(next_state, i, s2);
            obs[i] = ob; reward[i] = rw; flags[i] = fl;
            uint32_t w[4];
            state_hist_words(s2, w);
            hist_one<KIND>(hp0, true, w, sh, acc);
        }
    } else {
        const int64_t r_round = (n + 31) & ~(int64_t)31;
        for (int64_t i = tid; i < r_round; i += nthreads) {
            if (!table_ready) { mbar_wait(&bar, 0); table_ready = true; }
            if (i < n) {
                S s2; int32_t ob, fl; float rw;
                Env::step1(p, lut, load_state1(state, i, S()), action[i], seed, goff + (uint64_t)i, step_ctr, s2, ob, rw, fl);
                store_state1(next_state, i, s2);
                obs[i] = ob; reward[i] = rw; flags[i] = fl;
                uint32_t w[4];
                state_hist_words(s2, w);
                hist_one<KIND>(hp0, true, w, sh, acc);
            }
            if (++pending > 254) { hist_flush<NW>(acc, sh, n_bits, lane); pending = 0; }
        }
    }
    __syncwarp();
    hist_flush<NW>(acc, sh, n_bits, lane);
    if (!table_ready) mbar_wait(&bar, 0);      // a CTA must not exit while its bulk copy may still be in flight
    __syncthreads();
    hist_finish(sh, hist, bins, peers, world, rank, wait, hist_out);
}

// ============================================================================ host ===
namespace {

inline int device_sms() {
    static thread_local int cached_dev = -1, cached_sms = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
        cached_dev = dev; cached_sms = sms;
    }
    return cached_sms;
}

// Persistent grid: enough CTAs to fill every SM with the kernel's resident-CTA count
// (occupancy queried once per (kernel, CTA size, dynamic smem) and cached).
template <class K>
inline int grid_for(K kernel, int64_t n_threads_needed, int threads = POMDP_THREADS, size_t smem = 0) {
    struct Entry { const void* fn; int threads; size_t smem; int per_sm; };
    static thread_local Entry cache[48];
    static thread_local int n_cached = 0;
    int per_sm = 0;
    for (int i = 0; i < n_cached; ++i)
        if (cache[i].fn == (const void*)kernel && cache[i].threads == threads && cache[i].smem == smem)
            per_sm = cache[i].per_sm;
    if (per_sm == 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm <= 0)
            per_sm = 2;
        if (n_cached < 48) cache[n_cached++] = Entry{(const void*)kernel, threads, smem, per_sm};
    }
    int64_t need = (n_threads_needed + threads - 1) / threads;
    const int64_t cap = (int64_t)device_sms() * per_sm;
    if (need < 1) need = 1;
    if (need <= cap) return (int)need;
    // balanced persistent loop: every thread runs the same number of grid-stride iterations (a 2^20-env batch on 296
    // resident CTAs would otherwise give 73 % of the threads two iterations and the rest one)
    const int64_t iters = (need + cap - 1) / cap;
    return (int)((need + iters - 1) / iters);
}

inline int finish(const char* what) {
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return host::fail((int)e, "%s: %s", what, cudaGetErrorString(e));
    return 0;
}

inline bool aligned16(const void* a, const void* b, const void* c, const void* d, const void* e, const void* f) {
    return (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c | (uintptr_t)d | (uintptr_t)e | (uintptr_t)f) & 15) == 0;
}

template <class K>
inline int allow_smem(K kernel, size_t smem) {   // dynamic shared memory above 48 KB is opt-in per kernel
    if (smem <= 48 * 1024) return 0;
    const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    return e == cudaSuccess ? 0 : host::fail((int)e, "cudaFuncSetAttribute(smem=%zu): %s", smem, cudaGetErrorString(e));
}

inline bool pdl_enabled() {
    static const bool on = [] { const char* e = getenv("POMDP_B200_NO_PDL"); return !(e && e[0] == '1'); }();
    return on;
}

// Launch with the programmatic-stream-serialization attribute (see pdl_wait above); also valid
// under stream capture, where it becomes a programmatic dependency edge of the CUDA graph.
template <class... KArgs, class... Args>
inline void launch_pdl(void (*kernel)(KArgs...), int grid, int threads, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

template <class Env>
constexpr size_t param_table_smem() {
    if constexpr (Env::kParamTable) return Env::kParamTableBytes; else return 0;
}

template <class Env, bool kPacked = false>
int launch_step(const typename Env::Params& p, const void* d_table, uint32_t table_bytes, uint32_t smem_bytes,
                const int32_t* state,
                const int32_t* action, int32_t* next_state, int32_t* obs, float* reward, int32_t* flags, int64_t n,
                int64_t goff, uint64_t seed, uint32_t step_ctr, void* stream, const char* what) {
    if (kPacked) { reward = reinterpret_cast<float*>(obs); flags = obs; }   // one result stream; keeps the checks below uniform
    int rc = host::check_io(state, action, next_state, obs, reward, flags, n);
    if (rc) return rc;
    if (n == 0) return 0;
    if (goff < 0) return host::fail(POMDP_E_BADARG, "%s: global_offset is negative", what);
    if (Env::kTable && (!d_table || ((uintptr_t)d_table & 15)))
        return host::fail(POMDP_E_BADARG, "%s: d_table must be a 16-byte aligned device pointer", what);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = Env::kTable ? smem_bytes : param_table_smem<Env>();
    const PhiloxKey key = philox_key(seed);
    if (aligned16(state, action, next_state, obs, reward, flags) && (goff & 3) == 0) {
        auto k = pomdp_step_kernel<Env, true, kPacked>;
        if ((rc = allow_smem(k, smem))) return rc;
        const int grid = grid_for(k, (n + 3) >> 2, POMDP_STEP_THREADS, smem);
        launch_pdl(k, grid, POMDP_STEP_THREADS, smem, st, p, d_table, state, action, next_state, obs, reward, flags, n,
                   (uint64_t)goff, key, step_ctr, table_bytes);
    } else {
        auto k = pomdp_step_kernel<Env, false, kPacked>;
        if ((rc = allow_smem(k, smem))) return rc;
        const int grid = grid_for(k, n, POMDP_STEP_THREADS, smem);
        launch_pdl(k, grid, POMDP_STEP_THREADS, smem, st, p, d_table, state, action, next_state, obs, reward, flags, n,
                   (uint64_t)goff, key, step_ctr, table_bytes);
    }
    return finish(what);
}

// pomdp_E_step_hist: the step with the belief histogram of the next states in its epilogue (pomdp_step_hist_kernel)
template <class Env>
int launch_step_hist(const typename Env::Params& p, const void* d_table, uint32_t table_bytes, uint32_t smem_bytes,
                     const int32_t* state, const int32_t* action, int32_t* next_state, int32_t* obs, float* reward,
                     int32_t* flags, int64_t n, int64_t goff, uint64_t seed, uint32_t step_ctr, const PomdpHistSink* sink,
                     void* stream, const char* what) {
    int rc = host::check_io(state, action, next_state, obs, reward, flags, n);
    if (rc) return rc;
    if (goff < 0) return host::fail(POMDP_E_BADARG, "%s: global_offset is negative", what);
    if (!sink || !sink->scratch || ((uintptr_t)sink->scratch & 7) || ((uintptr_t)sink->hist_out & 7))
        return host::fail(POMDP_E_BADARG, "%s: the histogram sink needs an 8-byte aligned scratch (and hist_out)", what);
    if (sink->d_peer_bufs) {
        if (sink->world < 1 || sink->world > POMDP_HIST_MAX_RANKS || sink->rank < 0 || sink->rank >= sink->world)
            return host::fail(POMDP_E_BADARG, "%s: bad world size or rank in the histogram sink", what);
    } else if (!sink->hist_out) {
        return host::fail(POMDP_E_BADARG, "%s: a local histogram sink (no peer table) needs hist_out", what);
    }
    if (Env::kTable && (!d_table || ((uintptr_t)d_table & 15)))
        return host::fail(POMDP_E_BADARG, "%s: d_table must be a 16-byte aligned device pointer", what);
    const int bins = host::hist_bins(Env::kHistKind, Env::hist_p0(p), 0);
    if (bins <= 0 || bins > POMDP_HIST_MAX_BINS) return host::fail(POMDP_E_BADARG, "%s: bad histogram shape", what);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = Env::kTable ? smem_bytes : param_table_smem<Env>();
    const PhiloxKey key = philox_key(seed);
    unsigned long long* scratch = (unsigned long long*)sink->scratch;
    unsigned long long* const* peers = (unsigned long long* const*)sink->d_peer_bufs;
    unsigned long long* out = (unsigned long long*)sink->hist_out;
    // n == 0 still launches: an empty shard hands out zeros / must not leave its peers waiting
    if (n == 0 || (aligned16(state, action, next_state, obs, reward, flags) && (goff & 3) == 0)) {
        auto k = pomdp_step_hist_kernel<Env, true>;
        if ((rc = allow_smem(k, smem))) return rc;
        const int grid = grid_for(k, (n + 3) >> 2, POMDP_STEP_THREADS, smem);
        launch_pdl(k, grid, POMDP_STEP_THREADS, smem, st, p, d_table, state, action, next_state, obs, reward, flags, n,
                   (uint64_t)goff, key, step_ctr, table_bytes, scratch, bins, peers, (int)sink->world, (int)sink->rank,
                   (int)(sink->wait != 0), out);
    } else {
        auto k = pomdp_step_hist_kernel<Env, false>;
        if ((rc = allow_smem(k, smem))) return rc;
        const int grid = grid_for(k, n, POMDP_STEP_THREADS, smem);
        launch_pdl(k, grid, POMDP_STEP_THREADS, smem, st, p, d_table, state, action, next_state, obs, reward, flags, n,
                   (uint64_t)goff, key, step_ctr, table_bytes, scratch, bins, peers, (int)sink->world, (int)sink->rank,
                   (int)(sink->wait != 0), out);
    }
    return finish(what);
}

template <class Env>
int launch_reset(const typename Env::Params& p, int32_t* state, int32_t* obs, const uint8_t* mask, int64_t n,
                 int64_t goff, uint64_t seed, uint32_t step_ctr, void* stream, const char* what) {
    if (n < 0 || (n > 0 && !state)) return host::fail(POMDP_E_BADARG, "%s: bad n or NULL state", what);
    if (n == 0) return 0;
    if (goff < 0) return host::fail(POMDP_E_BADARG, "%s: global_offset is negative", what);
    if ((((uintptr_t)state) & 15) == 0 && (goff & 3) == 0 && !mask && obs && (((uintptr_t)obs) & 15) == 0) {
        auto k = pomdp_reset_kernel<Env, true, true>;
        const int grid = grid_for(k, (n + 3) >> 2);
        k<<<grid, POMDP_THREADS, 0, (cudaStream_t)stream>>>(p, state, obs, mask, n, (uint64_t)goff, philox_key(seed), step_ctr);
    } else if ((((uintptr_t)state) & 15) == 0 && (goff & 3) == 0) {
        auto k = pomdp_reset_kernel<Env, true>;
        const int grid = grid_for(k, (n + 3) >> 2);
        k<<<grid, POMDP_THREADS, 0, (cudaStream_t)stream>>>(p, state, obs, mask, n, (uint64_t)goff, philox_key(seed), step_ctr);
    } else {
        auto k = pomdp_reset_kernel<Env, false>;
        const int grid = grid_for(k, n);
        k<<<grid, POMDP_THREADS, 0, (cudaStream_t)stream>>>(p, state, obs, mask, n, (uint64_t)goff, philox_key(seed), step_ctr);
    }
    return finish(what);
}

template <class Env>
int launch_policy(const typename Env::Params& p, const void* d_table, uint32_t table_bytes, uint32_t smem_bytes,
                  const int32_t* state, int32_t* action, int64_t n, int64_t goff, uint64_t seed, uint32_t step_ctr,
                  void* stream, const char* what) {
    int rc = host::check_policy(state, action, n, goff, what);
    if (rc) return rc;
    if (n == 0) return 0;
    if (Env::kTable && (!d_table || ((uintptr_t)d_table & 15)))
        return host::fail(POMDP_E_BADARG, "%s: d_table must be a 16-byte aligned device pointer", what);
    const size_t smem = Env::kTable ? smem_bytes : 0;
    const PhiloxKey key = philox_key(seed);
    if ((((uintptr_t)state | (uintptr_t)action) & 15) == 0 && (goff & 3) == 0) {
        auto k = pomdp_policy_kernel<Env, true>;
        if ((rc = allow_smem(k, smem))) return rc;
        k<<<grid_for(k, (n + 3) >> 2, POMDP_THREADS, smem), POMDP_THREADS, smem, (cudaStream_t)stream>>>(
            p, d_table, state, action, n, (uint64_t)goff, key, step_ctr, table_bytes);
    } else {
        auto k = pomdp_policy_kernel<Env, false>;
        if ((rc = allow_smem(k, smem))) return rc;
        k<<<grid_for(k, n, POMDP_THREADS, smem), POMDP_THREADS, smem, (cudaStream_t)stream>>>(
            p, d_table, state, action, n, (uint64_t)goff, key, step_ctr, table_bytes);
    }
    return finish(what);
}

template <class Env>
int launch_rollout(const typename Env::Params& p, const void* d_table, uint32_t table_bytes, uint32_t smem_bytes,
                   const int32_t* state, const int32_t* first_action, int32_t* final_state, double* ret, int32_t* steps,
                   int32_t* flags, int64_t n,
                   int64_t goff, uint64_t seed, uint32_t step_ctr, int32_t max_steps, double discount, void* stream,
                   const char* what) {
    int rc = host::check_rollout(state, final_state, ret, steps, flags, n, goff, max_steps, what);
    if (rc) return rc;
    if ((uintptr_t)first_action & 3) return host::fail(POMDP_E_ALIGN, "%s: first_action must be 4-byte aligned", what);
    if (n == 0) return 0;
    if (Env::kTable && (!d_table || ((uintptr_t)d_table & 15)))
        return host::fail(POMDP_E_BADARG, "%s: d_table must be a 16-byte aligned device pointer", what);
    const size_t smem = Env::kTable ? smem_bytes : param_table_smem<Env>();
    const PhiloxKey key = philox_key(seed);
    const uintptr_t any = (uintptr_t)state | (uintptr_t)final_state | (uintptr_t)ret | (uintptr_t)steps | (uintptr_t)flags;
    if ((any & 15) == 0 && (goff & 3) == 0) {
        auto k = pomdp_rollout_kernel<Env, true>;
        if ((rc = allow_smem(k, smem))) return rc;
        k<<<grid_for(k, (n + 3) >> 2, POMDP_THREADS, smem), POMDP_THREADS, smem, (cudaStream_t)stream>>>(
            p, d_table, state, first_action, final_state, ret, steps, flags, n, (uint64_t)goff, key, step_ctr, max_steps,
            discount, table_bytes);
    } else {
        auto k = pomdp_rollout_kernel<Env, false>;
        if ((rc = allow_smem(k, smem))) return rc;
        k<<<grid_for(k, n, POMDP_THREADS, smem), POMDP_THREADS, smem, (cudaStream_t)stream>>>(
            p, d_table, state, first_action, final_state, ret, steps, flags, n, (uint64_t)goff, key, step_ctr, max_steps,
            discount, table_bytes);
    }
    return finish(what);
}

template <class Env>
int launch_obs_prob(const typename Env::Params& p, const void* d_table, uint32_t table_bytes, uint32_t smem_bytes,
                    const int32_t* state, const int32_t* action, const int32_t* obs, double* prob, int64_t n, double extra,
                    void* stream, const char* what) {
    int rc = host::check_obs_prob(state, action, obs, prob, n, what);
    if (rc) return rc;
    if (n == 0) return 0;
    if (Env::kTable && (!d_table || ((uintptr_t)d_table & 15)))
        return host::fail(POMDP_E_BADARG, "%s: d_table must be a 16-byte aligned device pointer", what);
    const size_t smem = Env::kTable ? smem_bytes : 0;
    auto k = pomdp_obs_prob_kernel<Env>;
    if ((rc = allow_smem(k, smem))) return rc;
    k<<<grid_for(k, n, POMDP_THREADS, smem), POMDP_THREADS, smem, (cudaStream_t)stream>>>(p, d_table, state, action, obs, prob, n,
                                                                                           extra, table_bytes);
    return finish(what);
}
template <class Env>
int launch_legal_mask(const typename Env::Params& p, const void* d_table, uint32_t table_bytes, uint32_t smem_bytes,
                      const int32_t* state, uint32_t* mask, int64_t n, void* stream, const char* what) {
    int rc = host::check_policy(state, mask, n, 0, what);
    if (rc) return rc;
    if (n == 0) return 0;
    if (Env::kTable && (!d_table || ((uintptr_t)d_table & 15)))
        return host::fail(POMDP_E_BADARG, "%s: d_table must be a 16-byte aligned device pointer", what);
    const size_t smem = Env::kTable ? smem_bytes : 0;
    auto k = pomdp_legal_mask_kernel<Env>;
    if ((rc = allow_smem(k, smem))) return rc;
    k<<<grid_for(k, n, POMDP_THREADS, smem), POMDP_THREADS, smem, (cudaStream_t)stream>>>(p, d_table, state, mask, n, table_bytes);
    return finish(what);
}

}  // namespace

extern "C" {

int pomdp_abi_version(void) { return POMDP_ABI_VERSION; }
const char* pomdp_last_error(void) { return host::err_buf(); }

// ---- Rock
int pomdp_rock_state_words(const PomdpRockParams* q) {
    int rc = host::make_rock(q, nullptr, nullptr);
    return rc ? rc : host::rock_words(q);
}
int64_t pomdp_rock_table_bytes(const PomdpRockParams* q) {
    int rc = host::make_rock(q, nullptr, nullptr);
    return rc ? (int64_t)rc : host::rock_table_bytes(q);
}
int pomdp_rock_build_table(const PomdpRockParams* q, void* host_table) {
    if (!host_table) return host::fail(POMDP_E_BADARG, "rock: host_table is NULL");
    RockDev d;
    return host::make_rock(q, &d, host_table);
}
int pomdp_rock_step(const PomdpRockParams* q, const void* d_table, const int32_t* state, const int32_t* action,
                    int32_t* next_state, int32_t* obs, float* reward, int32_t* flags, int64_t n, int64_t goff,
                    uint64_t seed, uint32_t step_ctr, void* stream) {
    RockDev d;
    int rc = host::make_rock(q, &d, nullptr);
    if (rc) return rc;
#define POMDP_ROCK_STEP(S, STOCH)                                                                                     \
    return launch_step<RockEnvT<S, STOCH>>(d, d_table, d.table_bytes, d.smem_bytes, state, action, next_state, obs,   \
                                           reward, flags, n, goff, seed, step_ctr, stream, "pomdp_rock_step")
    if (host::rock_words(q) == 1) {
        if (d.stochastic) POMDP_ROCK_STEP(uint32_t, true);
        POMDP_ROCK_STEP(uint32_t, false);
    }
    if (d.stochastic) POMDP_ROCK_STEP(uint64_t, true);
    POMDP_ROCK_STEP(uint64_t, false);
#undef POMDP_ROCK_STEP
}
int pomdp_rock_reset(const PomdpRockParams* q, const void* d_table, int32_t* state, int32_t* obs, const uint8_t* mask,
                     int64_t n, int64_t goff, uint64_t seed, uint32_t step_ctr, void* stream) {
    (void)d_table;
    RockDev d;
    int rc = host::make_rock(q, &d, nullptr);
    if (rc) return rc;
    if (host::rock_words(q) == 1)
        return launch_reset<RockEnvT<uint32_t, false>>(d, state, obs, mask, n, goff, seed, step_ctr, stream, "pomdp_rock_reset");
    return launch_reset<RockEnvT<uint64_t, false>>(d, state, obs, mask, n, goff, seed, step_ctr, stream, "pomdp_rock_reset");
}

// ---- Tag
int64_t pomdp_tag_table_bytes(void) { return (int64_t)sizeof(TagTables); }
int pomdp_tag_build_table(void* host_table) {
    if (!host_table) return host::fail(POMDP_E_BADARG, "tag: host_table is NULL");
    tag_build_tables((TagTables*)host_table);
    return 0;
}
int pomdp_tag_step(const PomdpTagParams* q, const void* d_table, const int32_t* state, const int32_t* action,
                   int32_t* next_state, int32_t* obs, float* reward, int32_t* flags, int64_t n, int64_t goff,
                   uint64_t seed, uint32_t step_ctr, void* stream) {
    TagDev d;
    int rc = host::make_tag(q, &d);
    if (rc) return rc;
    const uint32_t tb = (uint32_t)sizeof(TagTables), tbs = TAG_TABLES_BASE_BYTES;   // one opponent: + the transition LUT
    if (d.n_opp == 1)
        return launch_step<TagEnvT<1>>(d, d_table, tb, tb, state, action, next_state, obs, reward, flags, n, goff, seed,
                                       step_ctr, stream, "pomdp_tag_step");
    return launch_step<TagEnvT<4>>(d, d_table, tbs, tbs, state, action, next_state, obs, reward, flags, n, goff, seed, step_ctr,
                                   stream, "pomdp_tag_step");
}
int pomdp_tag_reset(const PomdpTagParams* q, int32_t* state, int32_t* obs, const uint8_t* mask, int64_t n, int64_t goff,
                    uint64_t seed, uint32_t step_ctr, void* stream) {
    TagDev d;
    int rc = host::make_tag(q, &d);
    if (rc) return rc;
    if (d.n_opp == 1)
        return launch_reset<TagEnvT<1>>(d, state, obs, mask, n, goff, seed, step_ctr, stream, "pomdp_tag_reset");
    return launch_reset<TagEnvT<4>>(d, state, obs, mask, n, goff, seed, step_ctr, stream, "pomdp_tag_reset");
}

// ---- Tiger
int pomdp_tiger_step(const PomdpTigerParams* q, const int32_t* state, const int32_t* action, int32_t* next_state,
                     int32_t* obs, float* reward, int32_t* flags, int64_t n, int64_t goff, uint64_t seed,
                     uint32_t step_ctr, void* stream) {
    TigerDev d;
    int rc = host::make_tiger(q, &d);
    if (rc) return rc;
    return launch_step<TigerEnvP>(d, nullptr, 0, 0, state, action, next_state, obs, reward, flags, n, goff, seed, step_ctr,
                                  stream, "pomdp_tiger_step");
}
int pomdp_tiger_reset(const PomdpTigerParams* q, int32_t* state, int32_t* obs, const uint8_t* mask, int64_t n,
                      int64_t goff, uint64_t seed, uint32_t step_ctr, void* stream) {
    TigerDev d;
    int rc = host::make_tiger(q, &d);
    if (rc) return rc;
    return launch_reset<TigerEnvP>(d, state, obs, mask, n, goff, seed, step_ctr, stream, "pomdp_tiger_reset");
}

// ---- Network
int pomdp_network_step(const PomdpNetworkParams* q, const int32_t* state, const int32_t* action, int32_t* next_state,
                       int32_t* obs, float* reward, int32_t* flags, int64_t n, int64_t goff, uint64_t seed,
                       uint32_t step_ctr, void* stream) {
    NetworkDev d;
    int rc = host::make_network(q, &d);
    if (rc) return rc;
    if (d.groups == 2)
        return launch_step<NetworkEnv10>(d, nullptr, 0, 0, state, action, next_state, obs, reward, flags, n, goff, seed, step_ctr,
                                         stream, "pomdp_network_step");
    return launch_step<NetworkEnvP>(d, nullptr, 0, 0, state, action, next_state, obs, reward, flags, n, goff, seed, step_ctr,
                                    stream, "pomdp_network_step");
}
int pomdp_network_reset(const PomdpNetworkParams* q, int32_t* state, int32_t* obs, const uint8_t* mask, int64_t n,
                        void* stream) {
    NetworkDev d;
    int rc = host::make_network(q, &d);
    if (rc) return rc;
    return launch_reset<NetworkEnvP>(d, state, obs, mask, n, 0, 0, 0, stream, "pomdp_network_reset");
}

// ---- BattleShip
int pomdp_battleship_step(const PomdpBattleshipParams* q, const int32_t* state, const int32_t* action,
                          int32_t* next_state, int32_t* obs, float* reward, int32_t* flags, int64_t n, void* stream) {
    ShipDev d;
    int rc = host::make_ship(q, &d);
    if (rc) return rc;
    rc = host::check_io(state, action, next_state, obs, reward, flags, n);
    if (rc) return rc;
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if ((((uintptr_t)state | (uintptr_t)next_state) & 15) == 0) {
        auto k = pomdp_battleship_step_kernel;
        const int grid = grid_for(k, n);
        k<<<grid, POMDP_THREADS, 0, st>>>(d, state, action, next_state, obs, reward, flags, n);
    } else {
        auto k = pomdp_battleship_step_plain_kernel;
        const int grid = grid_for(k, n);
        k<<<grid, POMDP_THREADS, 0, st>>>(d, state, action, next_state, obs, reward, flags, n);
    }
    return finish("pomdp_battleship_step");
}
int64_t pomdp_battleship_table_bytes(const PomdpBattleshipParams* q) { return host::make_ship_table(q, nullptr); }
int pomdp_battleship_build_table(const PomdpBattleshipParams* q, void* host_table) {
    if (!host_table) return host::fail(POMDP_E_BADARG, "battleship: host_table is NULL");
    const int64_t rc = host::make_ship_table(q, host_table);
    return rc < 0 ? (int)-rc : 0;
}
int pomdp_battleship_reset(const PomdpBattleshipParams* q, const void* d_table, int32_t* state, int32_t* obs, int32_t* flags,
                           const uint8_t* mask, int64_t n, int64_t goff, uint64_t seed, uint32_t step_ctr,
                           void* stream) {
    ShipDev d;
    int rc = d_table ? host::make_ship_tabled(q, &d) : host::make_ship(q, &d);
    if (rc) return rc;
    if (q->max_len - 1 > SHIP_MAX_SHIPS) return host::fail(POMDP_E_BADARG, "battleship: more than 8 ships");
    if (n < 0 || (n > 0 && !state)) return host::fail(POMDP_E_BADARG, "pomdp_battleship_reset: bad n or NULL state");
    if (goff < 0) return host::fail(POMDP_E_BADARG, "pomdp_battleship_reset: global_offset is negative");
    if ((uintptr_t)d_table & 15) return host::fail(POMDP_E_ALIGN, "pomdp_battleship_reset: d_table must be 16-byte aligned");
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (!d_table) {
        auto k = pomdp_battleship_reset_bitboard_kernel;
        k<<<grid_for(k, n), POMDP_THREADS, 0, st>>>(d, state, obs, flags, mask, n, (uint64_t)goff, philox_key(seed), step_ctr);
    } else {
        const bool bulk = !mask && (((uintptr_t)state) & 15) == 0;
        const bool lean = d.max_len <= 3 && d.tbl_n_tabled >= (uint32_t)(d.max_len - 1);
        auto k = bulk ? (lean ? pomdp_battleship_reset_table_kernel<true, true> : pomdp_battleship_reset_table_kernel<true, false>)
                      : (lean ? pomdp_battleship_reset_table_kernel<false, true> : pomdp_battleship_reset_table_kernel<false, false>);
        k<<<grid_for(k, n, POMDP_SHIP_RESET_THREADS), POMDP_SHIP_RESET_THREADS, 0, st>>>(
            d, (const unsigned char*)d_table, state, obs, flags, mask, n, (uint64_t)goff, philox_key(seed), step_ctr);
    }
    return finish("pomdp_battleship_reset");
}
int pomdp_battleship_reset_warpscan(const PomdpBattleshipParams* q, int32_t* state, int32_t* obs, int32_t* flags,
                                    const uint8_t* mask, int64_t n, int64_t goff, uint64_t seed, uint32_t step_ctr,
                                    void* stream) {
    ShipDev d;
    int rc = host::make_ship(q, &d);
    if (rc) return rc;
    if (q->max_len - 1 > SHIP_MAX_SHIPS) return host::fail(POMDP_E_BADARG, "battleship: more than 8 ships");
    if (n < 0 || (n > 0 && !state)) return host::fail(POMDP_E_BADARG, "pomdp_battleship_reset_warpscan: bad n or NULL state");
    if (n == 0) return 0;
    auto k = pomdp_battleship_reset_scan_kernel;
    const int grid = grid_for(k, n * 32);
    k<<<grid, POMDP_THREADS, 0, (cudaStream_t)stream>>>(d, state, obs, flags, mask, n, (uint64_t)goff, philox_key(seed), step_ctr);
    return finish("pomdp_battleship_reset_warpscan");
}
int pomdp_battleship_reset_rejection(const PomdpBattleshipParams* q, int32_t* state, int32_t* obs, int32_t* flags,
                                     const uint8_t* mask, int64_t n, int64_t goff, uint64_t seed, uint32_t step_ctr,
                                     void* stream) {
    ShipDev d;
    int rc = host::make_ship(q, &d);
    if (rc) return rc;
    if (n < 0 || (n > 0 && !state)) return host::fail(POMDP_E_BADARG, "pomdp_battleship_reset_rejection: bad n or NULL state");
    if (n == 0) return 0;
    auto k = pomdp_battleship_reset_rejection_kernel;
    const int grid = grid_for(k, n);
    k<<<grid, POMDP_THREADS, 0, (cudaStream_t)stream>>>(d, state, obs, flags, mask, n, (uint64_t)goff, philox_key(seed), step_ctr);
    return finish("pomdp_battleship_reset_rejection");
}

// ---- step with the compact result stream (obs | flags << 8 | reward_units << 16)
int pomdp_rock_step_packed(const PomdpRockParams* q, const void* d_table, const int32_t* state, const int32_t* action,
                           int32_t* next_state, int32_t* result, int64_t n, int64_t goff, uint64_t seed, uint32_t step_ctr,
                           void* stream) {
    RockDev d;
    int rc = host::make_rock(q, &d, nullptr);
    if (rc) return rc;
#define POMDP_ROCK_STEPP(S, STOCH)                                                                                    \
    return launch_step<RockEnvT<S, STOCH>, true>(d, d_table, d.table_bytes, d.smem_bytes, state, action, next_state,  \
                                                 result, nullptr, nullptr, n, goff, seed, step_ctr, stream,          \
                                                 "pomdp_rock_step_packed")
    if (host::rock_words(q) == 1) {
        if (d.stochastic) POMDP_ROCK_STEPP(uint32_t, true);
        POMDP_ROCK_STEPP(uint32_t, false);
    }
    if (d.stochastic) POMDP_ROCK_STEPP(uint64_t, true);
    POMDP_ROCK_STEPP(uint64_t, false);
#undef POMDP_ROCK_STEPP
}
int pomdp_tag_step_packed(const PomdpTagParams* q, const void* d_table, const int32_t* state, const int32_t* action,
                          int32_t* next_state, int32_t* result, int64_t n, int64_t goff, uint64_t seed, uint32_t step_ctr,
                          void* stream) {
    TagDev d;
    int rc = host::make_tag(q, &d);
    if (rc) return rc;
    const uint32_t tb = (uint32_t)sizeof(TagTables), tbs = TAG_TABLES_BASE_BYTES;   // one opponent: + the transition LUT
    if (d.n_opp == 1)
        return launch_step<TagEnvT<1>, true>(d, d_table, tb, tb, state, action, next_state, result, nullptr, nullptr, n, goff,
                                             seed, step_ctr, stream, "pomdp_tag_step_packed");
    return launch_step<TagEnvT<4>, true>(d, d_table, tbs, tbs, state, action, next_state, result, nullptr, nullptr, n, goff, seed,
                                         step_ctr, stream, "pomdp_tag_step_packed");
}
int pomdp_tiger_step_packed(const PomdpTigerParams* q, const int32_t* state, const int32_t* action, int32_t* next_state,
                            int32_t* result, int64_t n, int64_t goff, uint64_t seed, uint32_t step_ctr, void* stream) {
    TigerDev d;
    int rc = host::make_tiger(q, &d);
    if (rc) return rc;
    return launch_step<TigerEnvP, true>(d, nullptr, 0, 0, state, action, next_state, result, nullptr, nullptr, n, goff, seed,
                                        step_ctr, stream, "pomdp_tiger_step_packed");
}
int pomdp_network_step_packed(const PomdpNetworkParams* q, const int32_t* state, const int32_t* action, int32_t* next_state,
                              int32_t* result, int64_t n, int64_t goff, uint64_t seed, uint32_t step_ctr, void* stream) {
    NetworkDev d;
    int rc = host::make_network(q, &d);
    if (rc) return rc;
    if (d.groups == 2)
        return launch_step<NetworkEnv10, true>(d, nullptr, 0, 0, state, action, next_state, result, nullptr, nullptr, n, goff, seed,
                                               step_ctr, stream, "pomdp_network_step_packed");
    return launch_step<NetworkEnvP, true>(d, nullptr, 0, 0, state, action, next_state, result, nullptr, nullptr, n, goff, seed,
                                          step_ctr, stream, "pomdp_network_step_packed");
}

// ---- the same on HOST buffers: chunked H2D -> step -> D2H on the pipe's streams (include/pomdp_b200.h)
// Three streams -- one per engine: host-to-device copies, kernels, device-to-host copies -- linked by events, so each
// copy engine sees ONE ordered queue and is never held up by another chunk's kernel or by the opposite direction.
// Slot k = the staging buffers chunk c uses when c % n_slots == k; two events per slot guard their reuse (inputs may be
// overwritten once the slot's kernel has run, outputs once they have been copied out).  Measured against one stream per
// slot (scripts/exp_host_pipe2.cu, B200): 0.925 vs 0.97-0.99 ms per 2^22-env call at 4 chunks; every copy costs ~7 us
// on top of its bytes on this box, which is why smaller chunks lose (8 chunks: 1.01 ms) although they fill faster.
struct HostPipe {
    uint32_t magic;
    int words, n_slots, device;
    int64_t chunk;
    int32_t* slot[8];          // [state W*chunk | action chunk | next_state W*chunk | result chunk]
    cudaStream_t s_in, s_k, s_out;
    cudaEvent_t ev_in[8], ev_k[8], ev_out[8];
};
static const uint32_t kHostPipeMagic = 0x50495045u;

static void host_pipe_free(HostPipe* hp) {
    for (int k = 0; k < 8; ++k) {
        if (hp->slot[k]) cudaFree(hp->slot[k]);
        if (hp->ev_in[k]) cudaEventDestroy(hp->ev_in[k]);
        if (hp->ev_k[k]) cudaEventDestroy(hp->ev_k[k]);
        if (hp->ev_out[k]) cudaEventDestroy(hp->ev_out[k]);
    }
    if (hp->s_in) cudaStreamDestroy(hp->s_in);
    if (hp->s_k) cudaStreamDestroy(hp->s_k);
    if (hp->s_out) cudaStreamDestroy(hp->s_out);
    hp->magic = 0;
    delete hp;
}

int pomdp_host_pipe_create(int state_words, int64_t chunk_envs, int n_slots, void** pipe_out) {
    if (!pipe_out) return host::fail(POMDP_E_BADARG, "pomdp_host_pipe_create: pipe_out is NULL");
    *pipe_out = nullptr;
    if (state_words < 1 || state_words > SHIP_WORDS || chunk_envs < 4 || (chunk_envs & 3) || chunk_envs > (1ll << 28) ||
        n_slots < 1 || n_slots > 8)
        return host::fail(POMDP_E_BADARG, "pomdp_host_pipe_create: state_words %d, chunk_envs %lld (multiple of 4), n_slots %d (1..8)",
                          state_words, (long long)chunk_envs, n_slots);
    HostPipe* hp = new HostPipe();
    memset(hp, 0, sizeof(*hp));
    hp->magic = kHostPipeMagic; hp->words = state_words; hp->n_slots = n_slots; hp->chunk = chunk_envs;
    cudaError_t e = cudaGetDevice(&hp->device);
    const size_t bytes = (size_t)(2 * state_words + 2) * (size_t)chunk_envs * sizeof(int32_t);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&hp->s_in, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&hp->s_k, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&hp->s_out, cudaStreamNonBlocking);
    for (int k = 0; k < n_slots && e == cudaSuccess; ++k) {
        e = cudaMalloc((void**)&hp->slot[k], bytes);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&hp->ev_in[k], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&hp->ev_k[k], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&hp->ev_out[k], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) {
        host_pipe_free(hp);
        (void)cudaGetLastError();
        return host::fail((int)e, "pomdp_host_pipe_create: %s", cudaGetErrorString(e));
    }
    *pipe_out = hp;
    return 0;
}
int pomdp_host_pipe_destroy(void* pipe) {
    HostPipe* hp = (HostPipe*)pipe;
    if (!hp || hp->magic != kHostPipeMagic) return host::fail(POMDP_E_BADARG, "pomdp_host_pipe_destroy: not a pipe");
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(hp->device);
    cudaStreamSynchronize(hp->s_in); cudaStreamSynchronize(hp->s_k); cudaStreamSynchronize(hp->s_out);
    host_pipe_free(hp);
    cudaSetDevice(prev);
    return 0;
}
// The host buffers must be host-complete when the call is made (it does not order itself after work the caller queued
// on its own streams); the call returns when every result has landed.  Runs on the pipe's device whatever the caller's
// current device is (restored on return).
int pomdp_step_packed_host(void* pipe, int kind, const void* params, const void* d_table, const int32_t* h_state,
                           const int32_t* h_action, int32_t* h_next_state, int32_t* h_result, int64_t n, int64_t goff,
                           uint64_t seed, uint32_t step_ctr) {
    HostPipe* hp = (HostPipe*)pipe;
    int rc = host::check_host_step(hp && hp->magic == kHostPipeMagic, hp ? hp->words : 0, kind, params, h_state, h_action,
                                   h_next_state, h_result, n, goff);
    if (rc || n == 0) return rc;
    int prev_dev = 0;
    cudaError_t e = cudaGetDevice(&prev_dev);
    if (e == cudaSuccess && prev_dev != hp->device) e = cudaSetDevice(hp->device);
    const int W = hp->words, S = hp->n_slots;
    const int64_t C = hp->chunk;
    int ci = 0;
    for (int64_t lo = 0; lo < n && rc == 0 && e == cudaSuccess; lo += C, ++ci) {
        const int64_t m = n - lo < C ? n - lo : C;
        const int k = ci % S;
        int32_t* d_state = hp->slot[k];
        int32_t* d_action = d_state + (size_t)W * C;
        int32_t* d_next = d_action + C;
        int32_t* d_result = d_next + (size_t)W * C;
        if (ci >= S) e = cudaStreamWaitEvent(hp->s_in, hp->ev_k[k], 0);               // the slot's inputs have been consumed
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_state, h_state + (size_t)lo * W, (size_t)m * W * sizeof(int32_t), cudaMemcpyHostToDevice, hp->s_in);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_action, h_action + lo, (size_t)m * sizeof(int32_t), cudaMemcpyHostToDevice, hp->s_in);
        if (e == cudaSuccess) e = cudaEventRecord(hp->ev_in[k], hp->s_in);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(hp->s_k, hp->ev_in[k], 0);
        if (e == cudaSuccess && ci >= S) e = cudaStreamWaitEvent(hp->s_k, hp->ev_out[k], 0);   // the slot's outputs have been drained
        if (e != cudaSuccess) break;
        switch (kind) {
            case POMDP_KIND_ROCK: rc = pomdp_rock_step_packed((const PomdpRockParams*)params, d_table, d_state, d_action, d_next, d_result, m, goff + lo, seed, step_ctr, hp->s_k); break;
            case POMDP_KIND_TAG: rc = pomdp_tag_step_packed((const PomdpTagParams*)params, d_table, d_state, d_action, d_next, d_result, m, goff + lo, seed, step_ctr, hp->s_k); break;
            case POMDP_KIND_TIGER: rc = pomdp_tiger_step_packed((const PomdpTigerParams*)params, d_state, d_action, d_next, d_result, m, goff + lo, seed, step_ctr, hp->s_k); break;
            default: rc = pomdp_network_step_packed((const PomdpNetworkParams*)params, d_state, d_action, d_next, d_result, m, goff + lo, seed, step_ctr, hp->s_k); break;
        }
        if (rc) break;
        e = cudaEventRecord(hp->ev_k[k], hp->s_k);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(hp->s_out, hp->ev_k[k], 0);
        if (e == cudaSuccess) e = cudaMemcpyAsync(h_next_state + (size_t)lo * W, d_next, (size_t)m * W * sizeof(int32_t), cudaMemcpyDeviceToHost, hp->s_out);
        if (e == cudaSuccess) e = cudaMemcpyAsync(h_result + lo, d_result, (size_t)m * sizeof(int32_t), cudaMemcpyDeviceToHost, hp->s_out);
        if (e == cudaSuccess) e = cudaEventRecord(hp->ev_out[k], hp->s_out);
    }
    {   // drain even after an error: the slots are reused by the next call
        const cudaError_t e1 = cudaStreamSynchronize(hp->s_in), e2 = cudaStreamSynchronize(hp->s_k), e3 = cudaStreamSynchronize(hp->s_out);
        if (e == cudaSuccess) e = e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3);
    }
    if (prev_dev != hp->device) cudaSetDevice(prev_dev);
    if (rc) return rc;
    if (e != cudaSuccess) { (void)cudaGetLastError(); return host::fail((int)e, "pomdp_step_packed_host: %s", cudaGetErrorString(e)); }
    return 0;
}

// ---- uniform-legal policy and fused rollouts
#define POMDP_ROCK_DISPATCH(FN, ...)                                                          \
    do {                                                                                      \
        if (host::rock_words(q) == 1) {                                                       \
            if (d.stochastic) return FN<RockEnvT<uint32_t, true>>(__VA_ARGS__);               \
            return FN<RockEnvT<uint32_t, false>>(__VA_ARGS__);                                \
        }                                                                                     \
        if (d.stochastic) return FN<RockEnvT<uint64_t, true>>(__VA_ARGS__);                   \
        return FN<RockEnvT<uint64_t, false>>(__VA_ARGS__);                                    \
    } while (0)

int pomdp_rock_policy(const PomdpRockParams* q, const void* d_table, const int32_t* state, int32_t* action, int64_t n,
                      int64_t goff, uint64_t seed, uint32_t step_ctr, void* stream) {
    RockDev d;
    int rc = host::make_rock(q, &d, nullptr);
    if (rc) return rc;
    POMDP_ROCK_DISPATCH(launch_policy, d, d_table, d.table_bytes, d.smem_bytes, state, action, n, goff, seed, step_ctr,
                        stream, "pomdp_rock_policy");
}
int pomdp_rock_rollout(const PomdpRockParams* q, const void* d_table, const int32_t* state, const int32_t* first_action, int32_t* final_state,
                       double* ret, int32_t* steps, int32_t* flags, int64_t n, int64_t goff, uint64_t seed,
                       uint32_t step_ctr, int32_t max_steps, double discount, void* stream) {
    RockDev d;
    int rc = host::make_rock(q, &d, nullptr);
    if (rc) return rc;
    POMDP_ROCK_DISPATCH(launch_rollout, d, d_table, d.table_bytes, d.smem_bytes, state, first_action, final_state, ret, steps, flags, n,
                        goff, seed, step_ctr, max_steps, discount, stream, "pomdp_rock_rollout");
}
#undef POMDP_ROCK_DISPATCH

int pomdp_tag_policy(const PomdpTagParams* q, const void* d_table, const int32_t* state, int32_t* action, int64_t n,
                     int64_t goff, uint64_t seed, uint32_t step_ctr, void* stream) {
    TagDev d;
    int rc = host::make_tag(q, &d);
    if (rc) return rc;
    const uint32_t tbs = TAG_TABLES_BASE_BYTES;
    return launch_policy<TagEnvT<1>>(d, d_table, tbs, tbs, state, action, n, goff, seed, step_ctr, stream, "pomdp_tag_policy");
}
int pomdp_tag_rollout(const PomdpTagParams* q, const void* d_table, const int32_t* state, const int32_t* first_action, int32_t* final_state, double* ret,
                      int32_t* steps, int32_t* flags, int64_t n, int64_t goff, uint64_t seed, uint32_t step_ctr,
                      int32_t max_steps, double discount, void* stream) {
    TagDev d;
    int rc = host::make_tag(q, &d);
    if (rc) return rc;
    const uint32_t tb = (uint32_t)sizeof(TagTables), tbs = TAG_TABLES_BASE_BYTES;   // one opponent: + the transition LUT
    if (d.n_opp == 1)
        return launch_rollout<TagEnvT<1>>(d, d_table, tb, tb, state, first_action, final_state, ret, steps, flags, n, goff, seed, step_ctr,
                                          max_steps, discount, stream, "pomdp_tag_rollout");
    return launch_rollout<TagEnvT<4>>(d, d_table, tbs, tbs, state, first_action, final_state, ret, steps, flags, n, goff, seed, step_ctr,
                                      max_steps, discount, stream, "pomdp_tag_rollout");
}
int pomdp_tiger_policy(const PomdpTigerParams* q, const int32_t* state, int32_t* action, int64_t n, int64_t goff,
                       uint64_t seed, uint32_t step_ctr, void* stream) {
    TigerDev d;
    int rc = host::make_tiger(q, &d);
    if (rc) return rc;
    return launch_policy<TigerEnvP>(d, nullptr, 0, 0, state, action, n, goff, seed, step_ctr, stream, "pomdp_tiger_policy");
}
int pomdp_tiger_rollout(const PomdpTigerParams* q, const int32_t* state, const int32_t* first_action, int32_t* final_state, double* ret,
                        int32_t* steps, int32_t* flags, int64_t n, int64_t goff, uint64_t seed, uint32_t step_ctr,
                        int32_t max_steps, double discount, void* stream) {
    TigerDev d;
    int rc = host::make_tiger(q, &d);
    if (rc) return rc;
    return launch_rollout<TigerEnvP>(d, nullptr, 0, 0, state, first_action, final_state, ret, steps, flags, n, goff, seed, step_ctr,
                                     max_steps, discount, stream, "pomdp_tiger_rollout");
}
int pomdp_network_policy(const PomdpNetworkParams* q, const int32_t* state, int32_t* action, int64_t n, int64_t goff,
                         uint64_t seed, uint32_t step_ctr, void* stream) {
    NetworkDev d;
    int rc = host::make_network(q, &d);
    if (rc) return rc;
    return launch_policy<NetworkEnvP>(d, nullptr, 0, 0, state, action, n, goff, seed, step_ctr, stream, "pomdp_network_policy");
}
int pomdp_network_rollout(const PomdpNetworkParams* q, const int32_t* state, const int32_t* first_action, int32_t* final_state, double* ret,
                          int32_t* steps, int32_t* flags, int64_t n, int64_t goff, uint64_t seed, uint32_t step_ctr,
                          int32_t max_steps, double discount, void* stream) {
    NetworkDev d;
    int rc = host::make_network(q, &d);
    if (rc) return rc;
    if (d.groups == 2)
        return launch_rollout<NetworkEnv10>(d, nullptr, 0, 0, state, first_action, final_state, ret, steps, flags, n, goff, seed,
                                            step_ctr, max_steps, discount, stream, "pomdp_network_rollout");
    return launch_rollout<NetworkEnvP>(d, nullptr, 0, 0, state, first_action, final_state, ret, steps, flags, n, goff, seed, step_ctr,
                                       max_steps, discount, stream, "pomdp_network_rollout");
}
int pomdp_battleship_policy(const PomdpBattleshipParams* q, const int32_t* state, int32_t* action, int64_t n,
                            int64_t goff, uint64_t seed, uint32_t step_ctr, void* stream) {
    ShipDev d;
    int rc = host::make_ship(q, &d);
    if (rc) return rc;
    if ((rc = host::check_policy(state, action, n, goff, "pomdp_battleship_policy"))) return rc;
    if (n == 0) return 0;
    auto k = pomdp_battleship_policy_kernel;
    k<<<grid_for(k, n), POMDP_THREADS, 0, (cudaStream_t)stream>>>(d, state, action, n, (uint64_t)goff, philox_key(seed), step_ctr);
    return finish("pomdp_battleship_policy");
}
int pomdp_battleship_rollout(const PomdpBattleshipParams* q, const int32_t* state, const int32_t* first_action, int32_t* final_state, double* ret,
                             int32_t* steps, int32_t* flags, int64_t n, int64_t goff, uint64_t seed, uint32_t step_ctr,
                             int32_t max_steps, double discount, void* stream) {
    ShipDev d;
    int rc = host::make_ship(q, &d);
    if (rc) return rc;
    if ((rc = host::check_rollout(state, final_state, ret, steps, flags, n, goff, max_steps, "pomdp_battleship_rollout")))
        return rc;
    if (n == 0) return 0;
    auto k = pomdp_battleship_rollout_kernel;
    k<<<grid_for(k, n), POMDP_THREADS, 0, (cudaStream_t)stream>>>(d, state, first_action, final_state, ret, steps, flags, n,
                                                                (uint64_t)goff, philox_key(seed), step_ctr, max_steps, discount);
    return finish("pomdp_battleship_rollout");
}

// ---- observation likelihoods (_compute_prob) and legal-action masks (_generate_legal): SURVEY.md §8f ranks 2-3
#define POMDP_ROCK_DISPATCH2(FN, ...)                                                                          \
    do {                                                                                                       \
        if (host::rock_words(q) == 1) return FN<RockEnvT<uint32_t, false>>(__VA_ARGS__);                       \
        return FN<RockEnvT<uint64_t, false>>(__VA_ARGS__);                                                     \
    } while (0)
int pomdp_rock_obs_prob(const PomdpRockParams* q, const void* d_table, const int32_t* next_state, const int32_t* action,
                        const int32_t* obs, double* prob, int64_t n, void* stream) {
    RockDev d;
    int rc = host::make_rock(q, &d, nullptr);
    if (rc) return rc;
    POMDP_ROCK_DISPATCH2(launch_obs_prob, d, d_table, d.table_bytes, d.smem_bytes, next_state, action, obs, prob, n, 0.0, stream,
                         "pomdp_rock_obs_prob");
}
int pomdp_rock_legal_mask(const PomdpRockParams* q, const void* d_table, const int32_t* state, uint32_t* mask, int64_t n,
                          void* stream) {
    RockDev d;
    int rc = host::make_rock(q, &d, nullptr);
    if (rc) return rc;
    POMDP_ROCK_DISPATCH2(launch_legal_mask, d, d_table, d.table_bytes, d.smem_bytes, state, mask, n, stream, "pomdp_rock_legal_mask");
}
#undef POMDP_ROCK_DISPATCH2
int pomdp_rock_legal_list(const PomdpRockParams* q, const void* d_table, const int32_t* state, uint32_t* list, int64_t n,
                          void* stream) {
    RockDev d;
    int rc = host::make_rock(q, &d, nullptr);
    if (rc) return rc;
    if ((rc = host::check_policy(state, list, n, 0, "pomdp_rock_legal_list"))) return rc;
    if (n == 0) return 0;
    if (!d_table || ((uintptr_t)d_table & 15)) return host::fail(POMDP_E_BADARG, "pomdp_rock_legal_list: d_table must be a 16-byte aligned device pointer");
    const size_t smem = d.smem_bytes;
    if (host::rock_words(q) == 1) {
        auto k = pomdp_rock_legal_list_kernel<uint32_t>;
        if ((rc = allow_smem(k, smem))) return rc;
        k<<<grid_for(k, n, POMDP_THREADS, smem), POMDP_THREADS, smem, (cudaStream_t)stream>>>(d, d_table, state, list, n, d.table_bytes);
    } else {
        auto k = pomdp_rock_legal_list_kernel<uint64_t>;
        if ((rc = allow_smem(k, smem))) return rc;
        k<<<grid_for(k, n, POMDP_THREADS, smem), POMDP_THREADS, smem, (cudaStream_t)stream>>>(d, d_table, state, list, n, d.table_bytes);
    }
    return finish("pomdp_rock_legal_list");
}
int pomdp_tag_obs_prob(const PomdpTagParams* q, const int32_t* next_state, const int32_t* action, const int32_t* obs,
                       double* prob, int64_t n, void* stream) {
    TagDev d;
    int rc = host::make_tag(q, &d);
    if (rc) return rc;
    return launch_obs_prob<TagNoTable>(d, nullptr, 0, 0, next_state, action, obs, prob, n, 0.0, stream, "pomdp_tag_obs_prob");
}
int pomdp_tag_legal_mask(const PomdpTagParams* q, const int32_t* state, uint32_t* mask, int64_t n, void* stream) {
    TagDev d;
    int rc = host::make_tag(q, &d);
    if (rc) return rc;
    return launch_legal_mask<TagNoTable>(d, nullptr, 0, 0, state, mask, n, stream, "pomdp_tag_legal_mask");
}
int pomdp_tiger_obs_prob(const PomdpTigerParams* q, const int32_t* next_state, const int32_t* action, const int32_t* obs,
                         double* prob, int64_t n, double correct_prob, void* stream) {
    TigerDev d;
    int rc = host::make_tiger(q, &d);
    if (rc) return rc;
    return launch_obs_prob<TigerEnvP>(d, nullptr, 0, 0, next_state, action, obs, prob, n, correct_prob, stream, "pomdp_tiger_obs_prob");
}
int pomdp_tiger_legal_mask(const PomdpTigerParams* q, const int32_t* state, uint32_t* mask, int64_t n, void* stream) {
    TigerDev d;
    int rc = host::make_tiger(q, &d);
    if (rc) return rc;
    return launch_legal_mask<TigerEnvP>(d, nullptr, 0, 0, state, mask, n, stream, "pomdp_tiger_legal_mask");
}
int pomdp_network_obs_prob(const PomdpNetworkParams* q, const int32_t* next_state, const int32_t* action, const int32_t* obs,
                           double* prob, int64_t n, void* stream) {
    NetworkDev d;
    int rc = host::make_network(q, &d);
    if (rc) return rc;
    return launch_obs_prob<NetworkEnvP>(d, nullptr, 0, 0, next_state, action, obs, prob, n, 0.0, stream, "pomdp_network_obs_prob");
}
int pomdp_network_legal_mask(const PomdpNetworkParams* q, const int32_t* state, uint32_t* mask, int64_t n, void* stream) {
    NetworkDev d;
    int rc = host::make_network(q, &d);
    if (rc) return rc;
    return launch_legal_mask<NetworkEnvP>(d, nullptr, 0, 0, state, mask, n, stream, "pomdp_network_legal_mask");
}
int pomdp_battleship_obs_prob(const PomdpBattleshipParams* q, const int32_t* next_state, const int32_t* action,
                              const int32_t* obs, double* prob, int64_t n, void* stream) {
    ShipDev d;
    int rc = host::make_ship(q, &d);
    if (rc) return rc;
    if ((rc = host::check_obs_prob(next_state, action, obs, prob, n, "pomdp_battleship_obs_prob"))) return rc;
    if (n == 0) return 0;
    auto k = pomdp_battleship_obs_prob_kernel;
    k<<<grid_for(k, n), POMDP_THREADS, 0, (cudaStream_t)stream>>>(d, next_state, action, obs, prob, n);
    return finish("pomdp_battleship_obs_prob");
}
int pomdp_battleship_legal_mask(const PomdpBattleshipParams* q, const int32_t* state, uint32_t* mask, int64_t n, void* stream) {
    ShipDev d;
    int rc = host::make_ship(q, &d);
    if (rc) return rc;
    if ((rc = host::check_policy(state, mask, n, 0, "pomdp_battleship_legal_mask"))) return rc;
    if (n == 0) return 0;
    auto k = pomdp_battleship_legal_mask_kernel;
    k<<<grid_for(k, n), POMDP_THREADS, 0, (cudaStream_t)stream>>>(d, state, mask, n);
    return finish("pomdp_battleship_legal_mask");
}

int pomdp_rock_belief_update(const PomdpRockParams* q, const void* d_table, const int32_t* next_state, const int32_t* action,
                             const int32_t* obs, int32_t* count, int32_t* measured, double* lkv, double* lkw,
                             double* prob_valuable, int64_t n, void* stream) {
    RockDev d;
    int rc = host::make_rock(q, &d, nullptr);
    if (rc) return rc;
    if ((rc = host::check_belief(next_state, action, obs, count, measured, lkv, lkw, prob_valuable, n))) return rc;
    if (n == 0) return 0;
    if (!d_table || ((uintptr_t)d_table & 15))
        return host::fail(POMDP_E_BADARG, "pomdp_rock_belief_update: d_table must be a 16-byte aligned device pointer");
    const size_t smem = d.smem_bytes;
    if (host::rock_words(q) == 1) {
        auto k = pomdp_rock_belief_update_kernel<uint32_t>;
        if ((rc = allow_smem(k, smem))) return rc;
        k<<<grid_for(k, n, POMDP_THREADS, smem), POMDP_THREADS, smem, (cudaStream_t)stream>>>(
            d, d_table, next_state, action, obs, count, measured, lkv, lkw, prob_valuable, n, d.table_bytes);
    } else {
        auto k = pomdp_rock_belief_update_kernel<uint64_t>;
        if ((rc = allow_smem(k, smem))) return rc;
        k<<<grid_for(k, n, POMDP_THREADS, smem), POMDP_THREADS, smem, (cudaStream_t)stream>>>(
            d, d_table, next_state, action, obs, count, measured, lkv, lkw, prob_valuable, n, d.table_bytes);
    }
    return finish("pomdp_rock_belief_update");
}

// ---- heuristic action sets and rollouts
int pomdp_rock_history_update(const PomdpRockParams* q, const int32_t* obs_field, const int32_t* action,
                              const int32_t* next_obs_field, int32_t* check_totals, int64_t n, void* stream) {
    int rc = host::make_rock(q, nullptr, nullptr);
    if (rc) return rc;
    if (n < 0) return host::fail(POMDP_E_BADARG, "pomdp_rock_history_update: n is negative");
    if (n == 0) return 0;
    if (!obs_field || !action || !next_obs_field || !check_totals)
        return host::fail(POMDP_E_BADARG, "pomdp_rock_history_update: a required array pointer is NULL");
    if (((uintptr_t)obs_field | (uintptr_t)action | (uintptr_t)next_obs_field | (uintptr_t)check_totals) & 3)
        return host::fail(POMDP_E_ALIGN, "pomdp_rock_history_update: array pointers must be 4-byte aligned");
    auto k = pomdp_rock_history_update_kernel;
    k<<<grid_for(k, n), POMDP_THREADS, 0, (cudaStream_t)stream>>>(q->num_rocks, obs_field, action, next_obs_field, check_totals, n);
    return finish("pomdp_rock_history_update");
}
extern "C++" {
namespace {
template <bool kPolicy>
int launch_rock_preferred(const PomdpRockParams* q, const void* d_table, const int32_t* state, const int32_t* count,
                          const int32_t* measured, const double* pv, const int32_t* totals, int32_t* out, int64_t n, int64_t goff,
                          uint64_t seed, uint32_t step_ctr, void* stream, const char* what) {
    RockDev d;
    int rc = host::make_rock(q, &d, nullptr);
    if (rc) return rc;
    if ((rc = host::check_policy(state, out, n, goff, what))) return rc;
    if ((((uintptr_t)count | (uintptr_t)measured | (uintptr_t)totals) & 3) || ((uintptr_t)pv & 7))
        return host::fail(POMDP_E_ALIGN, "%s: int32 planes must be 4-byte and the float64 plane 8-byte aligned", what);
    if (n == 0) return 0;
    if (!d_table || ((uintptr_t)d_table & 15)) return host::fail(POMDP_E_BADARG, "%s: d_table must be a 16-byte aligned device pointer", what);
    const size_t smem = d.smem_bytes;
    const PhiloxKey key = philox_key(seed);
    if (host::rock_words(q) == 1) {
        auto k = pomdp_rock_preferred_kernel<uint32_t, kPolicy>;
        if ((rc = allow_smem(k, smem))) return rc;
        k<<<grid_for(k, n, POMDP_THREADS, smem), POMDP_THREADS, smem, (cudaStream_t)stream>>>(
            d, d_table, state, count, measured, pv, totals, out, n, (uint64_t)goff, key, step_ctr, d.table_bytes);
    } else {
        auto k = pomdp_rock_preferred_kernel<uint64_t, kPolicy>;
        if ((rc = allow_smem(k, smem))) return rc;
        k<<<grid_for(k, n, POMDP_THREADS, smem), POMDP_THREADS, smem, (cudaStream_t)stream>>>(
            d, d_table, state, count, measured, pv, totals, out, n, (uint64_t)goff, key, step_ctr, d.table_bytes);
    }
    return finish(what);
}
template <typename S, bool STOCH>
int launch_rock_rollout_preferred(const RockDev& d, const void* d_table, const int32_t* state, const int32_t* first_action,
                                  const RockPlanesPtr& pl, int32_t* final_state, double* ret, int32_t* steps, int32_t* flags,
                                  int64_t n, int64_t goff, uint64_t seed, uint32_t step_ctr, int32_t max_steps, double discount,
                                  int32_t next_is_reward, void* stream) {
    // the record flavour: no plane comes in or goes out, a scratch is there, and the 16-bit fields of a record cannot overflow
    const bool records = pl.scratch && !pl.count && !pl.measured && !pl.lkv && !pl.lkw && !pl.pv && !pl.totals && max_steps <= 32767;
    auto k = records ? pomdp_rock_rollout_preferred_kernel<S, STOCH, true> : pomdp_rock_rollout_preferred_kernel<S, STOCH, false>;
    int rc = allow_smem(k, d.smem_bytes);
    if (rc) return rc;
    k<<<grid_for(k, n, POMDP_THREADS, d.smem_bytes), POMDP_THREADS, d.smem_bytes, (cudaStream_t)stream>>>(
        d, d_table, state, first_action, pl, final_state, ret, steps, flags, n, (uint64_t)goff, philox_key(seed), step_ctr,
        max_steps, discount, next_is_reward, d.table_bytes);
    return finish("pomdp_rock_rollout_preferred");
}
}  // namespace
}  // extern "C++"
int pomdp_rock_preferred_mask(const PomdpRockParams* q, const void* d_table, const int32_t* state, const int32_t* count,
                              const int32_t* measured, const double* pv, const int32_t* totals, uint32_t* mask, int64_t n,
                              void* stream) {
    return launch_rock_preferred<false>(q, d_table, state, count, measured, pv, totals, (int32_t*)mask, n, 0, 0, 0, stream,
                                        "pomdp_rock_preferred_mask");
}
int pomdp_rock_policy_preferred(const PomdpRockParams* q, const void* d_table, const int32_t* state, const int32_t* count,
                                const int32_t* measured, const double* pv, const int32_t* totals, int32_t* action, int64_t n,
                                int64_t goff, uint64_t seed, uint32_t step_ctr, void* stream) {
    return launch_rock_preferred<true>(q, d_table, state, count, measured, pv, totals, action, n, goff, seed, step_ctr, stream,
                                       "pomdp_rock_policy_preferred");
}
int pomdp_rock_rollout_preferred(const PomdpRockParams* q, const void* d_table, const int32_t* state, const int32_t* first_action,
                                 const PomdpRockHeuristicPlanes* planes, int32_t* final_state, double* ret, int32_t* steps,
                                 int32_t* flags, int64_t n, int64_t goff, uint64_t seed, uint32_t step_ctr, int32_t max_steps,
                                 double discount, int32_t next_is_reward, void* stream) {
    const char* what = "pomdp_rock_rollout_preferred";
    RockDev d;
    int rc = host::make_rock(q, &d, nullptr);
    if (rc) return rc;
    if ((rc = host::check_rollout(state, final_state, ret, steps, flags, n, goff, max_steps, what))) return rc;
    if ((uintptr_t)first_action & 3) return host::fail(POMDP_E_ALIGN, "%s: first_action must be 4-byte aligned", what);
    RockPlanesPtr pl = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    if (planes) {
        pl.count = planes->count; pl.measured = planes->measured; pl.lkv = planes->lkv; pl.lkw = planes->lkw;
        pl.pv = planes->prob_valuable; pl.totals = planes->check_totals; pl.prev_obs = planes->prev_obs; pl.scratch = planes->scratch;
        if ((((uintptr_t)pl.count | (uintptr_t)pl.measured | (uintptr_t)pl.totals | (uintptr_t)pl.prev_obs) & 3) ||
            (((uintptr_t)pl.lkv | (uintptr_t)pl.lkw | (uintptr_t)pl.pv) & 7) || ((uintptr_t)pl.scratch & 31))
            return host::fail(POMDP_E_ALIGN, "%s: int32 planes must be 4-byte, float64 planes 8-byte and the scratch 32-byte aligned", what);
    }
    if (n == 0) return 0;
    if (!d_table || ((uintptr_t)d_table & 15)) return host::fail(POMDP_E_BADARG, "%s: d_table must be a 16-byte aligned device pointer", what);
#define POMDP_RRP(S, STOCH)                                                                                              \
    return launch_rock_rollout_preferred<S, STOCH>(d, d_table, state, first_action, pl, final_state, ret, steps, flags, n, goff, \
                                                   seed, step_ctr, max_steps, discount, next_is_reward, stream)
    if (host::rock_words(q) == 1) {
        if (d.stochastic) POMDP_RRP(uint32_t, true);
        POMDP_RRP(uint32_t, false);
    }
    if (d.stochastic) POMDP_RRP(uint64_t, true);
    POMDP_RRP(uint64_t, false);
#undef POMDP_RRP
}
extern "C++" {
namespace {
template <bool kPolicy>
int launch_tag_preferred(const PomdpTagParams* q, const void* d_table, const int32_t* state, const int32_t* last_obs,
                         const int32_t* last_action, int32_t* out, int64_t n, int64_t goff, uint64_t seed, uint32_t step_ctr,
                         void* stream, const char* what) {
    TagDev d;
    int rc = host::make_tag(q, &d);
    if (rc) return rc;
    if ((rc = host::check_policy(state, out, n, goff, what))) return rc;
    if (((uintptr_t)last_obs | (uintptr_t)last_action) & 3) return host::fail(POMDP_E_ALIGN, "%s: array pointers must be 4-byte aligned", what);
    if (n == 0) return 0;
    if (!d_table || ((uintptr_t)d_table & 15)) return host::fail(POMDP_E_BADARG, "%s: d_table must be a 16-byte aligned device pointer", what);
    const size_t smem = TAG_TABLES_BASE_BYTES;
    auto k = pomdp_tag_preferred_kernel<kPolicy>;
    k<<<grid_for(k, n, POMDP_THREADS, smem), POMDP_THREADS, smem, (cudaStream_t)stream>>>(d_table, state, last_obs, last_action, out, n,
                                                                                           (uint64_t)goff, philox_key(seed), step_ctr);
    return finish(what);
}
}  // namespace
}  // extern "C++"
int pomdp_tag_preferred_mask(const PomdpTagParams* q, const void* d_table, const int32_t* state, const int32_t* last_obs,
                             const int32_t* last_action, uint32_t* mask, int64_t n, void* stream) {
    return launch_tag_preferred<false>(q, d_table, state, last_obs, last_action, (int32_t*)mask, n, 0, 0, 0, stream,
                                       "pomdp_tag_preferred_mask");
}
int pomdp_tag_policy_preferred(const PomdpTagParams* q, const void* d_table, const int32_t* state, const int32_t* last_obs,
                               const int32_t* last_action, int32_t* action, int64_t n, int64_t goff, uint64_t seed,
                               uint32_t step_ctr, void* stream) {
    return launch_tag_preferred<true>(q, d_table, state, last_obs, last_action, action, n, goff, seed, step_ctr, stream,
                                      "pomdp_tag_policy_preferred");
}
int pomdp_tag_rollout_preferred(const PomdpTagParams* q, const void* d_table, const int32_t* state, int32_t* last_obs,
                                int32_t* last_action, const int32_t* first_action, int32_t* final_state, double* ret,
                                int32_t* steps, int32_t* flags, int64_t n, int64_t goff, uint64_t seed, uint32_t step_ctr,
                                int32_t max_steps, double discount, void* stream) {
    const char* what = "pomdp_tag_rollout_preferred";
    TagDev d;
    int rc = host::make_tag(q, &d);
    if (rc) return rc;
    if ((rc = host::check_rollout(state, final_state, ret, steps, flags, n, goff, max_steps, what))) return rc;
    if (((uintptr_t)first_action | (uintptr_t)last_obs | (uintptr_t)last_action) & 3)
        return host::fail(POMDP_E_ALIGN, "%s: array pointers must be 4-byte aligned", what);
    if (n == 0) return 0;
    if (!d_table || ((uintptr_t)d_table & 15)) return host::fail(POMDP_E_BADARG, "%s: d_table must be a 16-byte aligned device pointer", what);
    const size_t smem = d.n_opp == 1 ? sizeof(TagTables) : (size_t)TAG_TABLES_BASE_BYTES;     // one opponent: the step reads the LUT
    if (d.n_opp == 1) {
        auto k = pomdp_tag_rollout_preferred_kernel<1>;
        k<<<grid_for(k, n, POMDP_THREADS, smem), POMDP_THREADS, smem, (cudaStream_t)stream>>>(
            d, d_table, state, last_obs, last_action, first_action, final_state, ret, steps, flags, n, (uint64_t)goff, philox_key(seed),
            step_ctr, max_steps, discount);
    } else {
        auto k = pomdp_tag_rollout_preferred_kernel<4>;
        k<<<grid_for(k, n, POMDP_THREADS, smem), POMDP_THREADS, smem, (cudaStream_t)stream>>>(
            d, d_table, state, last_obs, last_action, first_action, final_state, ret, steps, flags, n, (uint64_t)goff, philox_key(seed),
            step_ctr, max_steps, discount);
    }
    return finish(what);
}

// ---- helpers
// Diagnostic (include/pomdp_b200.h): the compute-free probe of the step kernels' traffic pattern.
int pomdp_stream_probe(const int32_t* state, const int32_t* action, int32_t* next_state, int32_t* obs, float* reward,
                       int32_t* flags, int64_t n, void* stream) {
    int rc = host::check_io(state, action, next_state, obs, reward, flags, n);
    if (rc) return rc;
    if (n == 0) return 0;
    if (!aligned16(state, action, next_state, obs, reward, flags) || (n & 3))
        return host::fail(POMDP_E_ALIGN, "pomdp_stream_probe: arrays must be 16-byte aligned and n a multiple of 4");
    auto k = pomdp_stream_probe_kernel;
    const int grid = grid_for(k, n >> 2, POMDP_STEP_THREADS, 0);
    launch_pdl(k, grid, POMDP_STEP_THREADS, 0, (cudaStream_t)stream, state, action, next_state, obs, reward, flags, n);
    return finish("pomdp_stream_probe");
}

int pomdp_stream_probe_words(int32_t state_words, const int32_t* state, const int32_t* action, int32_t* next_state, int32_t* obs,
                             float* reward, int32_t* flags, int64_t n, void* stream) {
    if (state_words == 1) return pomdp_stream_probe(state, action, next_state, obs, reward, flags, n, stream);
    if (state_words != 2) return host::fail(POMDP_E_BADARG, "pomdp_stream_probe_words: state_words must be 1 or 2");
    int rc = host::check_io(state, action, next_state, obs, reward, flags, n);
    if (rc) return rc;
    if (n == 0) return 0;
    if (!aligned16(state, action, next_state, obs, reward, flags) || (n & 3))
        return host::fail(POMDP_E_ALIGN, "pomdp_stream_probe_words: arrays must be 16-byte aligned and n a multiple of 4");
    auto k = pomdp_stream_probe2_kernel;
    const int grid = grid_for(k, n >> 2, POMDP_STEP_THREADS, 0);
    launch_pdl(k, grid, POMDP_STEP_THREADS, 0, (cudaStream_t)stream, state, action, next_state, obs, reward, flags, n);
    return finish("pomdp_stream_probe_words");
}

int pomdp_coord_op(int32_t op, int32_t xs, int32_t ys, const int32_t* a, const int32_t* b, int32_t* out, int64_t n,
                   void* stream) {
    const int rc = host::check_coord_op(op, xs, a, b, out, n);
    if (rc) return rc;
    if (n == 0) return 0;
    auto k = pomdp_coord_kernel;
    const int grid = grid_for(k, n);
    k<<<grid, POMDP_THREADS, 0, (cudaStream_t)stream>>>(op, xs, ys, a, b, out, n);
    return finish("pomdp_coord_op");
}

int pomdp_belief_hist_bins(int32_t kind, int32_t p0, int32_t p1) { return host::hist_bins(kind, p0, p1); }
typedef void (*HistKernel)(int, int, const int32_t*, int, int64_t, unsigned long long*, int, unsigned long long* const*, int, int,
                           int, unsigned long long*);
// The carry-save flavour pays off only when every thread makes at least eight trips of four 16-byte loads.
static HistKernel hist_kernel_for(int32_t kind, int32_t words, int64_t n) {
    const int64_t groups = (words == 1 || words == 2) ? n >> (words == 1 ? 2 : 1) : 0;
    const bool csa = groups >= (int64_t)8 * 4 * 1024 * device_sms();
    switch (kind) {
        case POMDP_KIND_ROCK: return csa ? pomdp_belief_hist_kernel<POMDP_KIND_ROCK, true> : pomdp_belief_hist_kernel<POMDP_KIND_ROCK, false>;
        case POMDP_KIND_NETWORK: return csa ? pomdp_belief_hist_kernel<POMDP_KIND_NETWORK, true> : pomdp_belief_hist_kernel<POMDP_KIND_NETWORK, false>;
        case POMDP_KIND_TAG: return pomdp_belief_hist_kernel<POMDP_KIND_TAG, false>;
        case POMDP_KIND_TIGER: return pomdp_belief_hist_kernel<POMDP_KIND_TIGER, false>;
        default: return pomdp_belief_hist_kernel<POMDP_KIND_BATTLESHIP, false>;
    }
}
int pomdp_belief_hist(int32_t kind, int32_t p0, int32_t p1, const int32_t* state, int32_t words, int64_t n,
                      long long* hist, void* stream) {
    const int rc = host::check_hist(kind, p0, p1, state, words, n, hist, POMDP_HIST_MAX_BINS);
    if (rc) return rc;
    const int bins = host::hist_bins(kind, p0, p1);
    if (n == 0) return 0;
    auto k = hist_kernel_for(kind, words, n);
    // one 1024-thread CTA per SM: every CTA ends with one global atomic per non-empty bin, all CTAs on the same few
    // hundred addresses, so the CTA count (not the batch) sets that cost
    const int64_t want = (n + 4095) / 4096;
    const int grid = (int)(want < device_sms() ? (want < 1 ? 1 : want) : device_sms());
    k<<<grid, 1024, 0, (cudaStream_t)stream>>>(p0, p1, state, words, n, (unsigned long long*)hist, bins, nullptr, 0, 0, 0,
                                               nullptr);
    return finish("pomdp_belief_hist");
}
// One launch, no zero-fill before it: the CTA that takes the last ticket moves the counts from the scratch to hist_out.
int pomdp_belief_hist_once(int32_t kind, int32_t p0, int32_t p1, const int32_t* state, int32_t words, int64_t n,
                           long long* scratch, long long* hist_out, void* stream) {
    const int rc = host::check_hist(kind, p0, p1, state, words, n, scratch, POMDP_HIST_MAX_BINS);
    if (rc) return rc;
    if (!scratch) return host::fail(POMDP_E_BADARG, "pomdp_belief_hist_once: scratch is NULL");
    if (!hist_out || ((uintptr_t)hist_out & 7)) return host::fail(POMDP_E_BADARG, "pomdp_belief_hist_once: hist_out is NULL or not 8-byte aligned");
    const int bins = host::hist_bins(kind, p0, p1);
    auto k = hist_kernel_for(kind, words, n);
    // n == 0 still launches: hist_out must come back all zero
    const int64_t want = (n + 4095) / 4096;
    const int grid = (int)(want < device_sms() ? (want < 1 ? 1 : want) : device_sms());
    k<<<grid, 1024, 0, (cudaStream_t)stream>>>(p0, p1, state, words, n, (unsigned long long*)scratch, bins, nullptr, 0, 0, 0,
                                               (unsigned long long*)hist_out);
    return finish("pomdp_belief_hist_once");
}
// The histogram fused with its all-reduce over peer memory (include/pomdp_b200.h).
int pomdp_belief_hist_allreduce(int32_t kind, int32_t p0, int32_t p1, const int32_t* state, int32_t words, int64_t n,
                                long long* scratch, const void* const* d_peer_bufs, int32_t world, int32_t rank, int32_t wait,
                                long long* hist_out, void* stream) {
    const int rc = host::check_hist(kind, p0, p1, state, words, n, scratch, POMDP_HIST_MAX_BINS);
    if (rc) return rc;
    if (!scratch || !d_peer_bufs || world < 1 || world > POMDP_HIST_MAX_RANKS || rank < 0 || rank >= world || ((uintptr_t)hist_out & 7))
        return host::fail(POMDP_E_BADARG, "pomdp_belief_hist_allreduce: bad peer table, world size, rank or output");
    const int bins = host::hist_bins(kind, p0, p1);
    auto k = hist_kernel_for(kind, words, n);
    // n == 0 still launches: an empty shard contributes nothing but must not leave its peers waiting
    const int64_t want = (n + 4095) / 4096;
    const int grid = (int)(want < device_sms() ? (want < 1 ? 1 : want) : device_sms());
    k<<<grid, 1024, 0, (cudaStream_t)stream>>>(p0, p1, state, words, n, (unsigned long long*)scratch, bins,
                                               (unsigned long long* const*)d_peer_bufs, world, rank, wait != 0,
                                               (unsigned long long*)hist_out);
    return finish("pomdp_belief_hist_allreduce");
}

// ---- step + belief histogram of the next states in ONE kernel (include/pomdp_b200.h)
int pomdp_rock_step_hist(const PomdpRockParams* q, const void* d_table, const int32_t* state, const int32_t* action,
                         int32_t* next_state, int32_t* obs, float* reward, int32_t* flags, int64_t n, int64_t goff,
                         uint64_t seed, uint32_t step_ctr, const PomdpHistSink* sink, void* stream) {
    RockDev d;
    int rc = host::make_rock(q, &d, nullptr);
    if (rc) return rc;
#define POMDP_ROCK_STEP_HIST(S, STOCH)                                                                                \
    return launch_step_hist<RockEnvT<S, STOCH>>(d, d_table, d.table_bytes, d.smem_bytes, state, action, next_state,   \
                                                obs, reward, flags, n, goff, seed, step_ctr, sink, stream, "pomdp_rock_step_hist")
    if (host::rock_words(q) == 1) {
        if (d.stochastic) POMDP_ROCK_STEP_HIST(uint32_t, true);
        POMDP_ROCK_STEP_HIST(uint32_t, false);
    }
    if (d.stochastic) POMDP_ROCK_STEP_HIST(uint64_t, true);
    POMDP_ROCK_STEP_HIST(uint64_t, false);
#undef POMDP_ROCK_STEP_HIST
}
int pomdp_tag_step_hist(const PomdpTagParams* q, const void* d_table, const int32_t* state, const int32_t* action,
                        int32_t* next_state, int32_t* obs, float* reward, int32_t* flags, int64_t n, int64_t goff,
                        uint64_t seed, uint32_t step_ctr, const PomdpHistSink* sink, void* stream) {
    TagDev d;
    int rc = host::make_tag(q, &d);
    if (rc) return rc;
    const uint32_t tb = (uint32_t)sizeof(TagTables), tbs = TAG_TABLES_BASE_BYTES;
    if (d.n_opp == 1)
        return launch_step_hist<TagEnvT<1>>(d, d_table, tb, tb, state, action, next_state, obs, reward, flags, n, goff, seed,
                                            step_ctr, sink, stream, "pomdp_tag_step_hist");
    return launch_step_hist<TagEnvT<4>>(d, d_table, tbs, tbs, state, action, next_state, obs, reward, flags, n, goff, seed,
                                        step_ctr, sink, stream, "pomdp_tag_step_hist");
}
int pomdp_tiger_step_hist(const PomdpTigerParams* q, const int32_t* state, const int32_t* action, int32_t* next_state,
                          int32_t* obs, float* reward, int32_t* flags, int64_t n, int64_t goff, uint64_t seed,
                          uint32_t step_ctr, const PomdpHistSink* sink, void* stream) {
    TigerDev d;
    int rc = host::make_tiger(q, &d);
    if (rc) return rc;
    return launch_step_hist<TigerEnvP>(d, nullptr, 0, 0, state, action, next_state, obs, reward, flags, n, goff, seed, step_ctr,
                                       sink, stream, "pomdp_tiger_step_hist");
}
int pomdp_network_step_hist(const PomdpNetworkParams* q, const int32_t* state, const int32_t* action, int32_t* next_state,
                            int32_t* obs, float* reward, int32_t* flags, int64_t n, int64_t goff, uint64_t seed,
                            uint32_t step_ctr, const PomdpHistSink* sink, void* stream) {
    NetworkDev d;
    int rc = host::make_network(q, &d);
    if (rc) return rc;
    if (d.groups == 2)
        return launch_step_hist<NetworkEnv10>(d, nullptr, 0, 0, state, action, next_state, obs, reward, flags, n, goff, seed,
                                              step_ctr, sink, stream, "pomdp_network_step_hist");
    return launch_step_hist<NetworkEnvP>(d, nullptr, 0, 0, state, action, next_state, obs, reward, flags, n, goff, seed, step_ctr,
                                         sink, stream, "pomdp_network_step_hist");
}

}  // extern "C"
