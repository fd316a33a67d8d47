// raptor_b200/csrc/tc.cuh -- sm_100a tensor-core plumbing for the actor GEMMs: tcgen05.mma (kind::tf32) with operands in shared memory
// (canonical K-major, no-swizzle core-matrix layout), accumulators in tensor memory (TMEM), 1-D TMA bulk copy for the weight image,
// mbarrier completion.  Inline PTX only; descriptor bit layouts follow cute/arch/mma_sm100_desc.hpp (SmemDescriptor / InstrDescriptor).
#pragma once
#include <cstdint>

namespace b200l2f { namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p){ return (uint32_t)__cvta_generic_to_shared(p); }
// warp index in a form the compiler knows to be warp-uniform (value broadcast from lane 0): TMEM addresses derived from it stay in uniform
// registers instead of costing one R2UR per tcgen05.ld / tcgen05.st
__device__ __forceinline__ int uniform_warp_index(){ return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }

// ---- mbarrier ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count){
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init(){ asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes){
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// try_wait suspends the thread for a hardware-defined time slice; a bounded number of retries turns a lost completion into a trap
// (kernel error reported to the host) instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity){
    const uint32_t addr = smem_u32(bar);
    for(uint32_t spins = 0; ; spins++){
        uint32_t ok;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}" : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if(ok) return;
        if(spins > (1u << 22)) __trap();
    }
}

// ---- TMA: 1-D bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP) --------------------------------------
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar){
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---- TMA: 1-D bulk copy shared -> global (bulk async-group completion; SASS: UBLKCP.S2G).  dst, src and bytes are multiples of 16.
__device__ __forceinline__ void bulk_store(void* dst_gmem, const void* src_smem, uint32_t bytes){
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait_read(){ asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }   // the source may be rewritten
__device__ __forceinline__ void bulk_store_wait_all(){ asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }         // the writes are complete

// ---- proxy / tcgen05 fences -------------------------------------------------------------------------------------
__device__ __forceinline__ void fence_async_smem(){ asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }   // generic-proxy STS -> visible to the tensor core
__device__ __forceinline__ void tc_fence_before(){ asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after(){ asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMEM allocation (one warp) -----------------------------------------------------------------------------------
template <uint32_t COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem){
    static_assert(COLS >= 32 && COLS <= 512 && (COLS & (COLS - 1)) == 0, "TMEM columns: power of two in [32, 512]");
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr){
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}

// ---- descriptors ----------------------------------------------------------------------------------------------------
// shared-memory matrix descriptor, K-major, SWIZZLE_NONE ("interleave"): in 16-byte units ((8,n),2):((1,SBO),LBO)
//   8 rows x 16 B = one core matrix (128 B contiguous); SBO = byte stride between 8-row groups; LBO = byte stride between the two
//   16-byte K chunks one K=8 tf32 instruction consumes.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes){
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
    return d;                 // base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE
}
// descriptor of the same tile `byte_offset` further on (a multiple of 16): the start-address field is the low 14 bits and shared-memory addresses
// are < 256 KB, so the field cannot carry -- one add instead of rebuilding the descriptor
__device__ __forceinline__ uint64_t smem_desc_advance(uint64_t desc, uint32_t byte_offset){ return desc + (uint64_t)(byte_offset >> 4); }
// one lane of a fully active warp (elect.sync): the form the compiler turns into a predicated tcgen05 issue without a per-lane loop
__device__ __forceinline__ bool elect_one(){
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
// instruction descriptor, kind::tf32, fp32 accumulate, A and B K-major, dense
__host__ __device__ constexpr uint32_t make_idesc_tf32(uint32_t M, uint32_t N){
    return (1u << 4) | (2u << 7) | (2u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T   (M x N x 8, tf32 in, fp32 accumulate) -- issued by ONE thread
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate){
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// same with the A operand in tensor memory (lane = row, one 32-bit column per tf32 element, 8 columns per instruction)
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate){
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// all previously issued MMAs of this thread arrive on the mbarrier when complete (implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar){
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- TMEM -> registers: 32 lanes x 32 bit, N consecutive columns; thread t of warp w reads lane 32w + t ---------------------------
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float* v){
    uint32_t r0, r1, r2, r3;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(taddr));
    v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v){
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
#pragma unroll
    for(int i = 0; i < 8; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v){
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
#pragma unroll
    for(int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}
// registers -> TMEM: thread t of warp w writes N consecutive columns of lane 32w + t
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float* v){
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                   "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])) : "memory");
}
__device__ __forceinline__ void tmem_st_wait(){ asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait(){ asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- 3xTF32 operand split: x ~= hi + lo with hi, lo exactly representable in tf32 (10-bit mantissa; the tensor core ignores the low 13 bits)
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo){
    hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    lo = x - hi;   // exact; its own low bits are dropped by the tensor core (relative 2^-21 of x)
}

}}  // namespace b200l2f::tc
