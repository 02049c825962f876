// corpus_gen.cu -- bench/test utilities that run on the device: the counter-based synthetic corpus
// generator (bit-identical host twin: tests/corpus.py synth_file) and a read-only streaming probe
// that measures this GPU's HBM read roofline in the same run.
#include <cuda_runtime.h>
#include <cstdint>

#include "kernels.h"

namespace gscan {

__device__ __forceinline__ uint64_t mix64(uint64_t x)
{
	x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
	x ^= x >> 27; x *= 0x94D049BB133111EBull;
	x ^= x >> 31;
	return x;
}

__device__ __forceinline__ uint64_t synth_block(uint64_t base, uint64_t j)
{
	const uint64_t h1 = mix64(base + j);
	const uint64_t h2 = mix64(h1 ^ 0xA5A5A5A5A5A5A5A5ull);
	uint64_t out = 0;
#pragma unroll
	for (int k = 0; k < 8; k++) {
		const uint32_t b1 = (uint32_t)(h1 >> (8 * k)) & 255u, b2 = (uint32_t)(h2 >> (8 * k)) & 255u;
		const uint32_t c = b2 < 3u ? 10u : 0x20u + ((b1 * 95u) >> 8);
		out |= (uint64_t)c << (8 * k);
	}
	return out;
}

// file_len must be a multiple of 16; one thread writes 16 bytes
__global__ void k_synth(uint8_t *dptr, uint64_t seed, uint64_t first_file_id, uint64_t n_files, uint64_t file_len, uint64_t stride)
{
	const uint64_t per_file = file_len / 16;
	const uint64_t total = n_files * per_file;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
		const uint64_t f = i / per_file, c = i - f * per_file;
		const uint64_t base = seed * 0x9E3779B97F4A7C15ull + (first_file_id + f) * 0xD1B54A32D192ED03ull;
		const uint64_t lo = synth_block(base, 2 * c), hi = synth_block(base, 2 * c + 1);
		uint4 v;
		v.x = (uint32_t)lo; v.y = (uint32_t)(lo >> 32); v.z = (uint32_t)hi; v.w = (uint32_t)(hi >> 32);
		*reinterpret_cast<uint4 *>(dptr + f * stride + c * 16) = v;
	}
}

__global__ void k_plant(uint8_t *dptr, uint64_t seed, uint64_t first_file_id, uint64_t n_files, uint64_t file_len, uint64_t stride,
                        const uint8_t *needle, uint32_t needle_len, uint32_t every)
{
	const uint64_t f = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (f >= n_files) return;
	const uint64_t id = first_file_id + f;
	if (id % every != every / 2 || file_len <= needle_len) return;
	const uint64_t o = mix64(seed * 0x8CB92BA72F3D8DD7ull + id) % (file_len - needle_len);
	for (uint32_t k = 0; k < needle_len; k++) dptr[f * stride + o + k] = needle[k];
}

cudaError_t launch_synth_corpus(uint8_t *dptr, uint64_t seed, uint64_t first_file_id, uint64_t n_files, uint64_t file_len,
                                uint64_t stride, const uint8_t *d_needle, uint32_t needle_len, uint32_t needle_every,
                                cudaStream_t st)
{
	k_synth<<<148 * 16, 256, 0, st>>>(dptr, seed, first_file_id, n_files, file_len, stride);
	if (d_needle && needle_len && needle_every) {
		const unsigned nb = (unsigned)((n_files + 127) / 128);
		k_plant<<<nb, 128, 0, st>>>(dptr, seed, first_file_id, n_files, file_len, stride, d_needle, needle_len, needle_every);
	}
	return cudaGetLastError();
}

__global__ void k_fill_u32(uint32_t *p, size_t n, uint32_t v)
{
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

// plain stores instead of cudaMemset: a memset-to-zero region is held in the L2's zero-clear compression state
// and the first partial write into each block has to expand it (measured: 0.3 ms for 512 8-byte stores)
cudaError_t launch_fill_u32(uint32_t *p, size_t n, uint32_t v, cudaStream_t st)
{
	k_fill_u32<<<148 * 8, 256, 0, st>>>(p, n, v);
	return cudaGetLastError();
}

// read-only probe: 16-byte loads, 4 in flight per thread, xor-reduce
__global__ void __launch_bounds__(512) k_read_probe(const uint4 *p, uint64_t n16, unsigned long long *sum)
{
	uint32_t acc = 0;
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	for (; i + 3 * stride < n16; i += 4 * stride) {
		const uint4 a = __ldcs(p + i), b = __ldcs(p + i + stride), c = __ldcs(p + i + 2 * stride), d = __ldcs(p + i + 3 * stride);
		acc ^= a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w ^ c.x ^ c.y ^ c.z ^ c.w ^ d.x ^ d.y ^ d.z ^ d.w;
	}
	for (; i < n16; i += stride) {
		const uint4 a = __ldcs(p + i);
		acc ^= a.x ^ a.y ^ a.z ^ a.w;
	}
#pragma unroll
	for (int d = 16; d; d >>= 1) acc ^= __shfl_xor_sync(0xffffffffu, acc, d);
	if ((threadIdx.x & 31) == 0 && acc) atomicXor(sum, (unsigned long long)acc);
}

cudaError_t launch_read_probe(const void *dptr, uint64_t bytes, unsigned long long *d_sum, int grid, cudaStream_t st)
{
	k_read_probe<<<grid, 512, 0, st>>>(reinterpret_cast<const uint4 *>(dptr), bytes / 16, d_sum);
	return cudaGetLastError();
}

} // namespace gscan
