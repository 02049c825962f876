// scan_kernels.cu -- the hot path: hand-written sm_100a kernels that replace the match loop of
// /root/reference/src/grab.cc:175-213 (one pcre_exec per match) with one persistent streaming
// pass over every byte of a batch of scan units.
//
// Structure (one CTA per SM, persistent):
//   warp kConsumerWarps   producer: one lane walks this CTA's tiles and issues ONE 1-D TMA bulk
//                         copy (cp.async.bulk, SASS UBLKCP) per tile into a kStages-deep shared-memory
//                         ring, completion signalled through an mbarrier (complete_tx::bytes)
//   warps 0..15           consumers: each owns a contiguous slice of the tile, reads it with
//                         conflict-free 16-byte LDS (lane l <- bytes [16l,16l+16) of a 512-byte row),
//                         runs the SWAR filter (4 bytes per 32-bit op) and, on the rare flagged
//                         lane, verifies from shared memory.  Matches cross lane / warp / tile
//                         borders through the halo kept in front of and behind every tile.
// Output order: each warp appends its candidates in position order to a private scratch list and,
// at the end of its slice, reserves a contiguous range of the global candidate buffer with one
// atomicAdd and records (base, n) in the segment table.  Segment ids are position ordered, so the
// resolve pass needs no sort.
//
// HBM traffic per tile: len + pre + post bytes read once (pre+post <= 3% at the default geometry),
// 8 bytes of segment table written per 2 KiB scanned, candidates only where they exist.
#include <cuda_runtime.h>
#include <cstdint>

#include "device_types.h"
#include "swar.h"
#include "kernels.h"

namespace gscan {

// ------------------------------------------------------------------------------------------
// PTX wrappers: mbarrier + TMA bulk copy
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
	asm volatile(
	    "{\n"
	    ".reg .pred p;\n"
	    "WAIT_%=:\n"
	    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
	    "@p bra DONE_%=;\n"
	    "bra WAIT_%=;\n"
	    "DONE_%=:\n"
	    "}\n" ::"r"(smem_u32(bar)),
	    "r"(parity)
	    : "memory");
}
// global -> shared bulk copy executed by the TMA unit; bytes and both addresses multiples of 16
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
	                 smem_u32(smem_dst)),
	             "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
	             : "memory");
}

// ------------------------------------------------------------------------------------------
// shared-memory layout
// ------------------------------------------------------------------------------------------
struct __align__(16) StageCtl {
	uint64_t full[kStages];
	uint64_t empty[kStages];
	TileDesc desc[kStages];
};

constexpr size_t kSmemBytes = (size_t)kStages * kStageStride + sizeof(StageCtl) + 128;
size_t scan_smem_bytes() { return kSmemBytes; }

// what a consumer warp knows about its slice of the current tile
struct Slice {
	const uint8_t *tile; // shared-memory address of the tile's byte 0
	uint32_t off;        // offset of the tile in its unit
	uint32_t ulen;       // unit length
	uint32_t tile_len;
	uint32_t begin, niter; // slice = [begin, begin + niter*512) in tile coordinates
};

struct Emitter {
	Cand *scratch; // this warp's private list (global memory, L2 resident)
	uint32_t n;    // warp-uniform

	// mm: per-lane 16-bit mask of matched bytes of the lane's chunk; lens via callback
	template <class LenFn>
	__device__ __forceinline__ void emit(uint32_t mm, uint32_t pos0, LenFn len_of, uint32_t lane)
	{
		uint32_t cnt = __popc(mm), incl = cnt;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
			if ((int)lane >= d) incl += v;
		}
		uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
		uint32_t idx = n + incl - cnt;
		while (mm) {
			uint32_t b = __ffs(mm) - 1;
			mm &= mm - 1;
			Cand c;
			c.pos = pos0 + b;
			c.len = len_of(b);
			scratch[idx++] = c;
		}
		n += total;
	}

	__device__ __forceinline__ void flush(const ScanArgs &A, uint32_t seg, uint32_t lane)
	{
		uint32_t base = 0;
		if (n) {
			unsigned long long b64 = 0;
			if (lane == 0) b64 = atomicAdd(A.cursor, (unsigned long long)n);
			b64 = __shfl_sync(0xffffffffu, b64, 0);
			__syncwarp();
			if (b64 + n <= (unsigned long long)A.cand_cap) {
				base = (uint32_t)b64;
				for (uint32_t i = lane; i < n; i += 32) A.cand[base + i] = scratch[i];
			}
			// else: the host sees cursor > cand_cap, grows the buffer and re-runs the scan
		}
		if (lane == 0) A.segs[seg] = SegEntry{base, n};
		n = 0;
	}
};

// ------------------------------------------------------------------------------------------
// FIXED engine: alternation of fixed-length byte-class sequences
// ------------------------------------------------------------------------------------------
template <int D, int K>
struct FixedEngine {
	typedef FixedParams Params;

	// first sequence (in preference order) matching with its anchor byte at tile position p; 0: none
	static __device__ uint32_t verify(const FixedParams &P, const Slice &S, int p)
	{
		const int q = p - (int)P.anchor;              // match start in tile coordinates (may be < 0: pre-halo)
		const long long qu = (long long)S.off + q;    // ... in unit coordinates
		if (qu < 0) return 0;
		for (uint32_t s = 0; s < P.nseq; s++) {
			const uint32_t len = P.seq_len[s];
			if ((unsigned long long)qu + len > S.ulen) continue;
			const uint32_t *pp = P.seq_pos + P.seq_off[s];
			uint32_t i = 0;
			for (; i < len; i++) {
				const uint32_t e = pp[i];
				const uint32_t b = S.tile[q + (int)i];
				const uint32_t cls = e >> 16;
				bool ok;
				if (cls == 0xffffu) ok = (b & (e & 0xffu)) == ((e >> 8) & 0xffu);
				else ok = (P.cls_bm[cls * 8 + (b >> 5)] >> (b & 31)) & 1u;
				if (!ok) break;
			}
			if (i == len) return len;
		}
		return 0;
	}

	static __device__ __forceinline__ void run(const FixedParams &P, const Slice &S, Emitter &E, uint32_t lane)
	{
		for (uint32_t it = 0; it < S.niter; it++) {
			const uint32_t c0 = S.begin + it * 512 + lane * 16;
			const uint4 a = *reinterpret_cast<const uint4 *>(S.tile + c0);
			uint32_t w[5];
			w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
			w[4] = D ? *reinterpret_cast<const uint32_t *>(S.tile + c0 + 16) : 0u;
			uint32_t acc = 0;
#pragma unroll
			for (int j = 0; j < 4; j++) {
				const uint32_t s = D ? __funnelshift_r(w[j], w[j + 1], 8 * D) : w[j];
#pragma unroll
				for (int k = 0; k < K; k++) acc |= pair_test(w[j], s, P.m0[k], P.v0[k], P.m1[k], P.v1[k]);
			}
			acc &= kHigh;
			if (__any_sync(0xffffffffu, acc != 0)) {
				// rare path: exact verification of every flagged byte of this lane's chunk
				uint32_t mm = 0;
				if (acc) {
#pragma unroll
					for (int j = 0; j < 4; j++) {
						const uint32_t s = D ? __funnelshift_r(w[j], w[j + 1], 8 * D) : w[j];
						uint32_t f = 0;
#pragma unroll
						for (int k = 0; k < K; k++) f |= pair_test(w[j], s, P.m0[k], P.v0[k], P.m1[k], P.v1[k]);
						f &= kHigh;
						while (f) {
							const int byte = (__ffs(f) - 1) >> 3;
							f &= f - 1;
							const int p = (int)c0 + j * 4 + byte;
							if (p < (int)S.tile_len && verify(P, S, p)) mm |= 1u << (j * 4 + byte);
						}
					}
				}
				const uint32_t pos0 = S.off + c0 - P.anchor; // unit offset of a match anchored at chunk byte 0
				E.emit(mm, pos0, [&](uint32_t b) -> uint32_t {
					return P.uniform_len ? P.uniform_len : verify(P, S, (int)(c0 + b));
				}, lane);
			}
		}
	}
};

// ------------------------------------------------------------------------------------------
// RUN engine: one byte class repeated {n,}: one candidate per maximal run of >= n class bytes
// ------------------------------------------------------------------------------------------
template <int NLO, int NHI>
struct RunEngine {
	typedef RunParams Params;

	static __device__ __forceinline__ uint32_t class_flags(const RunParams &P, uint32_t x)
	{
		const uint32_t x7 = x & kLow7;
		uint32_t lo = 0, hi = 0;
#pragma unroll
		for (int r = 0; r < NLO; r++) lo |= range7(x7, P.add_ge_lo[r], P.add_gt_lo[r]);
#pragma unroll
		for (int r = 0; r < NHI; r++) hi |= range7(x7, P.add_ge_hi[r], P.add_gt_hi[r]);
		if (NHI) return ((lo & ~x) | (hi & x)) & kHigh;
		return lo & ~x & kHigh;
	}
	static __device__ __forceinline__ bool in_class(const RunParams &P, uint32_t b)
	{
		return (P.bitmap[b >> 5] >> (b & 31)) & 1u;
	}
	// 16-bit class mask of the 16 bytes at tile position c, bytes at or past the unit end read as 0
	static __device__ __forceinline__ uint32_t mask16(const RunParams &P, const Slice &S, uint32_t c)
	{
		const uint4 a = *reinterpret_cast<const uint4 *>(S.tile + c);
		uint32_t r = pack_top_nibble(class_flags(P, a.w)) >> 28;
		r = __funnelshift_l(pack_top_nibble(class_flags(P, a.z)), r, 4);
		r = __funnelshift_l(pack_top_nibble(class_flags(P, a.y)), r, 4);
		r = __funnelshift_l(pack_top_nibble(class_flags(P, a.x)), r, 4);
		const long long rem = (long long)S.ulen - (long long)S.off - (long long)c;
		const uint32_t valid = rem >= 16 ? 0xffffu : (rem <= 0 ? 0u : ((1u << (int)rem) - 1u));
		return r & valid;
	}

	static __device__ __forceinline__ void run(const RunParams &P, const Slice &S, Emitter &E, uint32_t lane)
	{
		if (S.niter == 0) return;
		const uint32_t nf = P.run_min < 17u ? P.run_min : 17u;
		// is the byte just before the slice in the class?  (the unit's first byte has no predecessor)
		uint32_t prevbit = 0;
		if (S.off + S.begin > 0) prevbit = in_class(P, S.tile[(int)S.begin - 1]);
		uint32_t cm_next = mask16(P, S, S.begin + lane * 16);
		for (uint32_t it = 0; it < S.niter; it++) {
			const uint32_t c0 = S.begin + it * 512 + lane * 16;
			const uint32_t cm = cm_next;
			if (it + 1 < S.niter) {
				cm_next = mask16(P, S, c0 + 512);
			} else {
				// the 16 bytes after this slice (next warp's slice, next tile via the post-halo, or past
				// the unit end): lanes 0..3 each classify one word, lane 0 assembles the mask
				const uint32_t cb = S.begin + S.niter * 512 + (lane & 3) * 4;
				const uint32_t x = *reinterpret_cast<const uint32_t *>(S.tile + cb);
				uint32_t nib = (pack_top_nibble(class_flags(P, x)) >> 28) << ((lane & 3) * 4);
				nib |= __shfl_xor_sync(0xffffffffu, nib, 1);
				nib |= __shfl_xor_sync(0xffffffffu, nib, 2);
				const long long rem = (long long)S.ulen - (long long)S.off - (long long)(S.begin + S.niter * 512);
				const uint32_t valid = rem >= 16 ? 0xffffu : (rem <= 0 ? 0u : ((1u << (int)rem) - 1u));
				cm_next = nib & valid;
			}
			const uint32_t la = __shfl_sync(0xffffffffu, cm_next, 0);
			uint32_t nx = __shfl_down_sync(0xffffffffu, cm, 1);
			if (lane == 31) nx = la;
			uint32_t pv = __shfl_up_sync(0xffffffffu, cm, 1) >> 15;
			if (lane == 0) pv = prevbit;
			prevbit = __shfl_sync(0xffffffffu, cm, 31) >> 15;
			const uint32_t starts = cm & ~((cm << 1) | pv);
			uint32_t cand = starts & runs_at_least(cm | (nx << 16), nf) & 0xffffu;
			if (__any_sync(0xffffffffu, cand != 0)) {
				if (P.run_min > 17u) {
					// long minimum: confirm bytes 17..n-1 from shared memory (post-halo covers n)
					uint32_t keep = 0;
					uint32_t t = cand;
					while (t) {
						const uint32_t b = __ffs(t) - 1;
						t &= t - 1;
						const uint32_t p = c0 + b;
						bool ok = (unsigned long long)S.off + p + P.run_min <= S.ulen;
						for (uint32_t i = 17; ok && i < P.run_min; i++) ok = in_class(P, S.tile[p + i]);
						if (ok) keep |= 1u << b;
					}
					cand = keep;
				}
				E.emit(cand, S.off + c0, [](uint32_t) -> uint32_t { return 0u; }, lane);
			}
		}
	}
};

// ------------------------------------------------------------------------------------------
// the persistent kernel
// ------------------------------------------------------------------------------------------
template <class Eng>
__global__ void __launch_bounds__(kScanThreads, 1) scan_kernel(const __grid_constant__ ScanArgs A, const __grid_constant__ typename Eng::Params P)
{
	extern __shared__ __align__(128) uint8_t smem[];
	StageCtl *ctl = reinterpret_cast<StageCtl *>(smem + (size_t)kStages * kStageStride);

	const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	if (threadIdx.x == 0) {
		for (int s = 0; s < kStages; s++) {
			mbar_init(&ctl->full[s], 1);
			mbar_init(&ctl->empty[s], kConsumerWarps);
		}
		mbar_fence_init();
	}
	__syncthreads();

	if (warp == kConsumerWarps) {
		// ---------------- producer ----------------
		if (lane == 0) {
			uint32_t stage = 0, phase = 0;
			for (uint32_t t = blockIdx.x; t < A.n_tiles; t += gridDim.x) {
				mbar_wait(&ctl->empty[stage], phase ^ 1u);
				const TileDesc d = A.tiles[t];
				ctl->desc[stage] = d;
				const uint32_t pre = d.off ? A.pre : 0u;
				uint32_t body = (d.ulen - d.off + 15u) & ~15u; // bytes of the unit from the tile start, padded to 16
				const uint32_t want = (uint32_t)kTileBytes + A.post;
				if (body > want) body = want;
				const uint32_t bytes = pre + body;
				uint8_t *dst = smem + (size_t)stage * kStageStride + kPreMax - pre;
				mbar_arrive_expect_tx(&ctl->full[stage], bytes);
				tma_load_1d(dst, reinterpret_cast<const void *>(d.src - pre), bytes, &ctl->full[stage]);
				if (++stage == kStages) { stage = 0; phase ^= 1u; }
			}
		}
		return;
	}

	// ---------------- consumers ----------------
	Emitter E;
	E.scratch = A.scratch + ((size_t)blockIdx.x * kConsumerWarps + warp) * kSubTileMax;
	E.n = 0;
	uint32_t stage = 0, phase = 0;
	for (uint32_t t = blockIdx.x; t < A.n_tiles; t += gridDim.x) {
		mbar_wait(&ctl->full[stage], phase);
		const TileDesc d = ctl->desc[stage];
		Slice S;
		S.tile = smem + (size_t)stage * kStageStride + kPreMax;
		S.off = d.off;
		S.ulen = d.ulen;
		S.tile_len = d.len;
		// slice length: a multiple of 512 so that every warp row is 32 x 16 contiguous bytes
		const uint32_t sub = ((d.len + kConsumerWarps - 1) / kConsumerWarps + 511u) & ~511u;
		S.begin = warp * sub;
		S.niter = 0;
		if (S.begin < d.len) {
			const uint32_t end = S.begin + sub < d.len ? S.begin + sub : d.len;
			S.niter = (end - S.begin + 511u) >> 9;
		}
		Eng::run(P, S, E, lane);
		E.flush(A, t * kConsumerWarps + warp, lane);
		__syncwarp();
		if (lane == 0) mbar_arrive(&ctl->empty[stage]);
		if (++stage == kStages) { stage = 0; phase ^= 1u; }
	}
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
template <class Eng>
static cudaError_t launch(const ScanArgs &A, const typename Eng::Params &P, int grid, cudaStream_t st)
{
	static bool configured = false; // per instantiation; attribute is per-device but idempotent
	cudaError_t e = cudaFuncSetAttribute(scan_kernel<Eng>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
	if (e != cudaSuccess) return e;
	configured = true;
	(void)configured;
	scan_kernel<Eng><<<grid, kScanThreads, kSmemBytes, st>>>(A, P);
	return cudaGetLastError();
}

template <int D>
static cudaError_t launch_fixed_d(const ScanArgs &A, const FixedParams &P, int grid, cudaStream_t st)
{
	switch (P.ntests) {
	case 1: return launch<FixedEngine<D, 1>>(A, P, grid, st);
	case 2: return launch<FixedEngine<D, 2>>(A, P, grid, st);
	case 3: return launch<FixedEngine<D, 3>>(A, P, grid, st);
	case 4: return launch<FixedEngine<D, 4>>(A, P, grid, st);
	case 5: case 6: return launch<FixedEngine<D, 6>>(A, P, grid, st);
	default: return launch<FixedEngine<D, 8>>(A, P, grid, st);
	}
}

cudaError_t launch_scan_fixed(const ScanArgs &A, const FixedParams &P, int delta, int grid, cudaStream_t st)
{
	switch (delta) {
	case 0: return launch_fixed_d<0>(A, P, grid, st);
	case 1: return launch_fixed_d<1>(A, P, grid, st);
	case 2: return launch_fixed_d<2>(A, P, grid, st);
	default: return launch_fixed_d<3>(A, P, grid, st);
	}
}

template <int NHI>
static cudaError_t launch_run_h(const ScanArgs &A, const RunParams &P, int grid, cudaStream_t st)
{
	switch (P.nlo) {
	case 0: case 1: return launch<RunEngine<1, NHI>>(A, P, grid, st);
	case 2: return launch<RunEngine<2, NHI>>(A, P, grid, st);
	case 3: return launch<RunEngine<3, NHI>>(A, P, grid, st);
	case 4: return launch<RunEngine<4, NHI>>(A, P, grid, st);
	case 5: case 6: return launch<RunEngine<6, NHI>>(A, P, grid, st);
	default: return launch<RunEngine<8, NHI>>(A, P, grid, st);
	}
}

cudaError_t launch_scan_run(const ScanArgs &A, const RunParams &P, int grid, cudaStream_t st)
{
	switch (P.nhi) {
	case 0: return launch_run_h<0>(A, P, grid, st);
	case 1: return launch_run_h<1>(A, P, grid, st);
	default: return launch_run_h<2>(A, P, grid, st);
	}
}

} // namespace gscan
