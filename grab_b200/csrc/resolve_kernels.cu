// resolve_kernels.cu -- turns the scan kernel's position-ordered segments into exactly the match
// sequence of the reference loop (/root/reference/src/grab.cc:175-213), entirely on the device:
//   1. exclusive prefix sum over the segment counts (three small kernels, no sort)
//   2. gather: candidates -> one array ordered by (unit, position)
//   3. select: per unit, replay the loop's control flow -- guard `start + minlen < end` (:175, quirk
//      Q1), leftmost candidate at or after `start` (:178), `start += ovector[1] + a` (:209), stop
//      after the first match in FIRST mode (:204-212), skip the rest of the line (<= 511 bytes) in
//      LINE mode (:194-196)
//   4. compact the kept records.
// Traffic: 8 bytes per 2 KiB scanned for the segment table plus O(candidates).
#include <cuda_runtime.h>
#include <cstdint>

#include "../../include/gscan.h"
#include "device_types.h"
#include "kernels.h"

namespace gscan {

constexpr int kScanBlock = 256;   // threads
constexpr int kPerThread = 8;     // items per thread
constexpr int kPerBlock = kScanBlock * kPerThread;

// block-wide exclusive scan of one value per thread; returns the block total via *total
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *total)
{
	__shared__ uint32_t warp_sums[kScanBlock / 32];
	__shared__ uint32_t block_total;
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t incl = v;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
		if ((int)lane >= d) incl += t;
	}
	if (lane == 31) warp_sums[warp] = incl;
	__syncthreads();
	if (warp == 0) {
		uint32_t w = lane < kScanBlock / 32 ? warp_sums[lane] : 0u;
		uint32_t wi = w;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			uint32_t t = __shfl_up_sync(0xffffffffu, wi, d);
			if ((int)lane >= d) wi += t;
		}
		if (lane < kScanBlock / 32) warp_sums[lane] = wi - w;
		if (lane == 31) block_total = wi;
	}
	__syncthreads();
	const uint32_t r = incl - v + warp_sums[warp];
	*total = block_total;
	__syncthreads();
	return r;
}

// ---- 1. segment counts -> block sums ----
__global__ void k_seg_sums(const SegEntry *segs, uint32_t n_segs, uint32_t tag, uint32_t *blk)
{
	const uint32_t base = blockIdx.x * kPerBlock + threadIdx.x * kPerThread;
	uint32_t s = 0;
#pragma unroll
	for (int i = 0; i < kPerThread; i++)
		if (base + i < n_segs) s += seg_count(segs[base + i], tag);
	uint32_t total;
	block_exclusive_scan(s, &total);
	if (threadIdx.x == 0) blk[blockIdx.x] = total;
}

// single block: exclusive scan of blk[0..nb) in place, grand total to *total_out (and extra_out)
__global__ void k_scan_blk(uint32_t *blk, uint32_t nb, uint32_t *total_out, uint32_t *extra_out)
{
	__shared__ uint32_t carry;
	if (threadIdx.x == 0) carry = 0;
	__syncthreads();
	for (uint32_t base = 0; base < nb; base += kScanBlock) {
		const uint32_t i = base + threadIdx.x;
		const uint32_t v = i < nb ? blk[i] : 0u;
		uint32_t total;
		const uint32_t ex = block_exclusive_scan(v, &total);
		const uint32_t c = carry;
		if (i < nb) blk[i] = ex + c;
		__syncthreads();
		if (threadIdx.x == 0) carry = c + total;
		__syncthreads();
	}
	if (threadIdx.x == 0) {
		*total_out = carry;
		if (extra_out) *extra_out = carry;
	}
}

// ---- 2. gather into (unit, pos) order ----
__global__ void k_gather(const ResolveArgs R)
{
	const uint32_t base = blockIdx.x * kPerBlock + threadIdx.x * kPerThread;
	uint32_t cnt[kPerThread];
	uint32_t s = 0;
#pragma unroll
	for (int i = 0; i < kPerThread; i++) {
		cnt[i] = base + i < R.n_segs ? seg_count(R.segs[base + i], R.tag) : 0u;
		s += cnt[i];
	}
	uint32_t total;
	uint32_t p = block_exclusive_scan(s, &total) + R.blk[blockIdx.x];
#pragma unroll
	for (int i = 0; i < kPerThread; i++) {
		const uint32_t seg = base + i;
		if (seg >= R.n_segs) break;
		const uint32_t tile = seg / R.slices_per_tile;
		const TileDesc td = R.tiles[tile];
		if (td.off == 0 && seg % R.slices_per_tile == 0) R.unit_start[td.unit] = p; // first segment of a unit
		if (cnt[i]) {
			const Cand *src = R.cand + R.segs[seg].base;
			for (uint32_t k = 0; k < cnt[i]; k++) {
				OutRec o;
				o.unit = td.unit; o.pos = src[k].pos; o.len = src[k].len; o.pad = 0;
				R.ord[p + k] = o;
			}
			p += cnt[i];
		}
	}
}

// ---- 3. per-unit replay of the reference loop ----
// Pass 1 (WRITE=false) counts the matches of every unit, pass 2 writes them at the unit's slot of
// the output (exclusive scan of the counts in between): the output can hold matches that are not
// candidates -- in LINE mode the RUN engine restarts up to 511 bytes after a match, possibly in
// the middle of a run, and PCRE then reports a match at that very byte.
__device__ __forceinline__ bool in_class(const ResolveArgs &R, uint32_t b) { return (R.bitmap[b >> 5] >> (b & 31)) & 1u; }

template <bool WRITE>
__global__ void k_walk(const ResolveArgs R)
{
	const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
	if (u >= R.n_units) return;
	uint32_t i = R.unit_start[u];
	const uint32_t end = R.unit_start[u + 1];
	uint32_t n = 0;
	FinalRec *o = WRITE ? R.out + R.unit_out[u] : nullptr;
	if (i != end) {
		const DevUnit du = R.units[u];
		const uint8_t *data = reinterpret_cast<const uint8_t *>(du.ptr);
		const uint64_t ulen = du.len;
		const bool run = R.engine == GSCAN_ENGINE_RUN;
		uint64_t start = 0;
		for (;;) {
			if (!(start + R.minlen < ulen)) break;                 // grab.cc:175 (strict '<': Q1)
			uint64_t pos = 0, e = 0;
			bool found = false;
			if (run && R.mode == GSCAN_MODE_LINE && start > 0 && in_class(R, data[start - 1]) && in_class(R, data[start])) {
				// `start` lies inside a run: the leftmost match, if the rest of the run is long enough, is AT start
				e = start;
				while (e < ulen && in_class(R, data[e])) e++;
				if (e - start >= R.run_min) { pos = start; found = true; }
			}
			if (!found) {
				while (i < end && R.ord[i].pos < start) i++;       // candidates swallowed by the previous match / skip
				if (i == end) break;
				pos = R.ord[i].pos;
				e = pos + R.ord[i].len;
				if (run) {                                         // greedy: extend to the end of the run
					e = pos + R.run_min;
					while (e < ulen && in_class(R, data[e])) e++;
				}
				i++;
			}
			if (WRITE) { FinalRec r; r.start = du.base_off + pos; r.file_id = du.file_id; r.len = (uint32_t)(e - pos); o[n] = r; }
			n++;
			if (R.mode == GSCAN_MODE_FIRST) break;                 // grab.cc:206 / :211
			if (R.mode == GSCAN_MODE_LINE) {                       // grab.cc:194-196: a = bytes to '\n', <= 511
				uint32_t a = 0;
				while (e + a < ulen && a < 511 && data[e + a] != '\n') a++;
				e += a;
			}
			start = e;                                             // grab.cc:209
		}
	}
	if (!WRITE) R.unit_out[u] = n;
}

// ---- 4. generic u32 block sums / exclusive scan (per-unit match counts -> output slots) ----
__global__ void k_u32_sums(const uint32_t *v, uint32_t n, uint32_t *blk)
{
	const uint32_t base = blockIdx.x * kPerBlock + threadIdx.x * kPerThread;
	uint32_t s = 0;
#pragma unroll
	for (int i = 0; i < kPerThread; i++)
		if (base + i < n) s += v[base + i];
	uint32_t total;
	block_exclusive_scan(s, &total);
	if (threadIdx.x == 0) blk[blockIdx.x] = total;
}

__global__ void k_u32_exclusive(uint32_t *v, uint32_t n, const uint32_t *blk)
{
	const uint32_t base = blockIdx.x * kPerBlock + threadIdx.x * kPerThread;
	uint32_t x[kPerThread];
	uint32_t s = 0;
#pragma unroll
	for (int i = 0; i < kPerThread; i++) { x[i] = base + i < n ? v[base + i] : 0u; s += x[i]; }
	uint32_t total;
	uint32_t p = block_exclusive_scan(s, &total) + blk[blockIdx.x];
#pragma unroll
	for (int i = 0; i < kPerThread; i++)
		if (base + i < n) { v[base + i] = p; p += x[i]; }
}

cudaError_t launch_resolve_count(const ResolveArgs &R, cudaStream_t st, uint32_t *launches)
{
	uint32_t nl = 0;
	const uint32_t nb_seg = (R.n_segs + kPerBlock - 1) / kPerBlock;
	k_seg_sums<<<nb_seg, kScanBlock, 0, st>>>(R.segs, R.n_segs, R.tag, R.blk); nl++;
	k_scan_blk<<<1, kScanBlock, 0, st>>>(R.blk, nb_seg, R.totals, R.unit_start + R.n_units); nl++;
	k_gather<<<nb_seg, kScanBlock, 0, st>>>(R); nl++;
	k_walk<false><<<(R.n_units + 127) / 128, 128, 0, st>>>(R); nl++;
	const uint32_t nb_u = (R.n_units + kPerBlock - 1) / kPerBlock;
	uint32_t *blk2 = R.blk + nb_seg + 1; // block sums of the per-unit counts live behind the segment block sums
	k_u32_sums<<<nb_u, kScanBlock, 0, st>>>(R.unit_out, R.n_units, blk2); nl++;
	k_scan_blk<<<1, kScanBlock, 0, st>>>(blk2, nb_u, R.totals + 1, nullptr); nl++;
	k_u32_exclusive<<<nb_u, kScanBlock, 0, st>>>(R.unit_out, R.n_units, blk2); nl++;
	if (launches) *launches = nl;
	return cudaGetLastError();
}

cudaError_t launch_resolve_write(const ResolveArgs &R, cudaStream_t st, uint32_t *launches)
{
	k_walk<true><<<(R.n_units + 127) / 128, 128, 0, st>>>(R);
	if (launches) *launches = 1;
	return cudaGetLastError();
}

} // namespace gscan
