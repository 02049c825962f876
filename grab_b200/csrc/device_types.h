// device_types.h -- PODs shared by host code and kernels (layout in HBM).
#pragma once
#include <cstdint>

namespace gscan {

// Geometry of the persistent scan kernel.
constexpr int kTileBytes = 32768;   // bytes of one tile: the unit of the TMA pipeline
constexpr int kConsumerWarps = 16;  // warps that scan; +1 producer warp issues the bulk copies
constexpr int kStages = 4;          // smem ring depth
constexpr int kPreMax = 256;        // bytes kept in front of a tile (filter anchor look-behind)
constexpr int kPostMax = 1056;      // bytes kept behind a tile (verification look-ahead)
constexpr int kStageBytes = kPreMax + kTileBytes + kPostMax; // 34080, multiple of 16
constexpr int kStageStride = ((kStageBytes + 127) / 128) * 128;
constexpr int kSubTileMax = ((kTileBytes / kConsumerWarps + 511) / 512) * 512; // bytes one warp owns per tile
constexpr int kScanThreads = (kConsumerWarps + 1) * 32;

// One tile of one unit.  Built on the host when a batch is planned, 32 bytes.
struct TileDesc {
	uint64_t src;   // global address of the tile's first byte (16-byte aligned)
	uint32_t unit;  // index into the batch's unit table
	uint32_t off;   // offset of the tile inside the unit
	uint32_t len;   // bytes of the unit inside this tile (<= kTileBytes)
	uint32_t ulen;  // length of the unit
	uint32_t pad[2];
};
static_assert(sizeof(TileDesc) == 32, "TileDesc layout");

// One segment == the part of one tile one consumer warp owns; candidates of a segment are
// contiguous and ordered in the candidate buffer.  seg id = tile * kConsumerWarps + warp.
struct SegEntry { uint32_t base, n; };

// A candidate / match inside a unit.
struct Cand { uint32_t pos, len; };

// An ordered, resolved match: what goes back to the host.
struct OutRec { uint32_t unit, pos, len, pad; };

struct FixedParams {
	uint32_t ntests;
	uint32_t m0[8], v0[8], m1[8], v1[8]; // byte replicated x4
	uint32_t anchor;      // pattern byte the filter stream is anchored on
	uint32_t nseq;
	uint32_t uniform_len; // != 0: every sequence has this length
	uint32_t maxlen;
	const uint16_t *seq_len;  // [nseq]
	const uint32_t *seq_off;  // [nseq] index into seq_pos
	const uint32_t *seq_pos;  // per position: mask | val << 8 | cls << 16 (cls 0xffff: none)
	const uint32_t *cls_bm;   // [ncls][8] bitmaps for positions that are not a masked equality
};

struct RunParams {
	uint32_t nlo, nhi;
	uint32_t add_ge_lo[8], add_gt_lo[8]; // (0x80-lo)*0x01010101, (0x7f-hi)*0x01010101 for ranges in 0x00-0x7f
	uint32_t add_ge_hi[2], add_gt_hi[2]; // same for ranges in 0x80-0xff (after clearing bit 7)
	uint32_t run_min;
	uint32_t bitmap[8];
};

struct ScanArgs {
	const TileDesc *tiles;
	uint32_t n_tiles;
	uint32_t pre;   // bytes to load in front of a tile that is not the first of its unit (multiple of 16)
	uint32_t post;  // bytes to load behind a tile (multiple of 16)
	Cand *cand;
	uint32_t cand_cap;
	unsigned long long *cursor; // [0]: candidates reserved so far
	SegEntry *segs;
	Cand *scratch;  // [gridDim.x * kConsumerWarps][kSubTileMax]
};

} // namespace gscan
