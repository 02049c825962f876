// kernels.h -- host-callable launchers of the CUDA kernels (scan_kernels.cu, resolve_kernels.cu,
// corpus_gen.cu).  Internal to libgscan.so; the public surface is include/gscan.h.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "device_types.h"

namespace gscan {

size_t scan_smem_bytes(const ScanGeom &g, bool look_behind);
ScanGeom scan_geom(int engine, uint32_t n_tests_or_ranges);
cudaError_t launch_scan_fixed(const ScanArgs &A, const FixedParams &P, int delta, const ScanGeom &g, int grid, cudaStream_t st);
cudaError_t launch_scan_hash(const ScanArgs &A, const HashParams &P, int grid, cudaStream_t st);
cudaError_t launch_scan_null(const ScanArgs &A, int geom, int grid, cudaStream_t st);
cudaError_t launch_scan_run(const ScanArgs &A, const RunParams &P, const ScanGeom &g, int grid, cudaStream_t st);

struct DevUnit {
	uint64_t ptr;        // device address of the unit's bytes
	uint64_t base_off;   // file offset of the unit (grab.cc:154 `off`)
	uint32_t len;
	uint32_t first_tile; // index of the unit's first tile
	uint32_t file_id;
	uint32_t pad;
};

// what goes back to the caller: byte-for-byte a gscan_match (include/gscan.h)
struct FinalRec {
	uint64_t start;      // base_off + pos == off + (start - content) + ovector[0]   (grab.cc:186)
	uint32_t file_id;
	uint32_t len;
};

struct ResolveArgs {
	const TileDesc *tiles;
	const SegEntry *segs;
	uint32_t n_segs;
	uint32_t spt_shift;   // log2(slices per tile)
	uint32_t tag;         // generation of the scan that filled segs
	const Cand *cand;
	const DevUnit *units;
	uint32_t n_units;
	OutRec *ord;          // candidates in (unit, pos) order
	FinalRec *out;        // selected matches, same order (capacity known after the count pass)
	uint32_t *unit_start; // [n_units + 1] first candidate of each unit in ord
	uint32_t *unit_out;   // [n_units] matches per unit, then (exclusive scan) first output slot
	uint32_t *blk;        // block-sum scratch
	uint32_t *totals;     // [0] candidates, [1] matches, [2] VM limit flag
	uint32_t mode, minlen, engine, run_min;
	uint32_t bitmap[8];   // RUN class, for match-length extension
	uint32_t total_cand;  // known on the host after the scan kernel
	const uint32_t *vm_code; // general patterns: VM program (3 words per instruction) and byte classes (8 words each)
	const uint32_t *vm_sets;
	uint32_t vm_runstart; // candidates are run starts of `bitmap`: also try the search start itself when it lies inside a run
	// chain path (FIXED patterns whose matches can overlap, or LINE mode, with many candidates per unit): the selected
	// matches of a unit are the chain c0 -> next(c0) -> ..., next(i) = first candidate at or after where the search resumes
	// behind i.  Pointer doubling marks every chain member in O(log n) passes over all candidates instead of one thread
	// walking a huge unit alone.  chain_buf: chain_levels x chain_cap jump tables, then the mark and rank arrays.
	uint32_t chain, chain_levels, chain_cap;
	uint32_t *chain_buf;
	// general patterns on the chain path (vm_par): patterns whose attempts do not depend on where the subject begins
	// (Program::vm_start_free) with many candidates per unit.  One anchored VM attempt per candidate, all at once
	// (k_vm_attempts: outcome and match length into ord[].pad / .len), the candidates that matched are compacted into
	// vm_ord / vm_unit_start (exclusive scan of vm_flag), and the chain kernels run over those: the selected matches of a
	// unit are the chain from its first matching candidate, next(i) = first matching candidate at or after where the loop
	// resumes behind i.  A candidate whose match has a capturing group set (Q2) or that ran into the VM's limits ends its
	// unit's chain there and is not reported (grab.cc:179: rc <= 0 => break).
	uint32_t vm_par;
	OutRec *vm_ord;           // [total_cand] the candidates with pad != 0, (unit, pos) order
	uint32_t *vm_unit_start;  // [n_units + 1] into vm_ord
	uint32_t *vm_flag;        // [total_cand] pad != 0, then its exclusive scan
	unsigned long long *vm_budget; // [n_units] VM steps the unit's attempts may still take, all of them together
	// dense general patterns on the chain path (vm_ready): start-free programs without a candidate filter, units big enough
	// that one thread per unit would walk for seconds.  k_vm_dense runs the anchored attempts of 64 consecutive positions
	// per thread over the whole batch, twice like the serial walk (count, prefix sum, write): what comes out IS the compacted
	// list of vm_par (vm_ord / vm_unit_start, totals[4]), and the chain kernels take it from there.
	uint32_t vm_ready;        // vm_ord / vm_unit_start are filled already: no segments, no gather, no k_vm_attempts
	uint32_t dense_blocks;    // 64-position blocks of the batch = n_tiles << (dense_tile_shift - 6)
	uint32_t dense_tile_shift;
	uint32_t vm_dense;    // general pattern without a candidate filter: no candidate list, the walk offers every position whose byte
	                      // is in `bitmap` (the first-byte set) to the VM
	uint32_t flat;        // ALL mode, RUN or a FIXED pattern whose matches can never overlap, no VM: every candidate of a unit is a
	                      // match, except that the loop guard (grab.cc:175, Q1) can drop the unit's LAST one -- no replay needed, the
	                      // write pass runs one thread per candidate (a 1 GiB window with 10^7 candidates resolves in microseconds)
};
// count pass: segment scan, gather, per-unit replay that counts, slot scan; totals[1] = number of matches
cudaError_t launch_resolve_count(const ResolveArgs &R, cudaStream_t st, uint32_t *launches); // vm_dense: no segments, no gather
// write pass: per-unit replay that writes R.out
cudaError_t launch_resolve_write(const ResolveArgs &R, cudaStream_t st, uint32_t *launches);
// dense general patterns, parallel: count pass (matching positions per 64-position block into vm_flag, exclusive scan, total
// into totals[4]) and, once the host has sized vm_ord, the write pass (+ vm_unit_start)
cudaError_t launch_vm_dense_count(const ResolveArgs &R, cudaStream_t st, uint32_t *launches);
cudaError_t launch_vm_dense_write(const ResolveArgs &R, cudaStream_t st, uint32_t *launches);

cudaError_t launch_synth_corpus(uint8_t *dptr, uint64_t seed, uint64_t first_file_id, uint64_t n_files, uint64_t file_len,
                                uint64_t stride, const uint8_t *d_needle, uint32_t needle_len, uint32_t needle_every,
                                cudaStream_t st);
cudaError_t launch_fill_u32(uint32_t *p, size_t n, uint32_t v, cudaStream_t st);
cudaError_t launch_read_probe(const void *dptr, uint64_t bytes, unsigned long long *d_sum, int grid, cudaStream_t st);

} // namespace gscan
