// Brute-force check of the SWAR primitives in grab_b200/csrc/swar.h, compiled for the HOST by tests/test_swar.py.
// (The product only ever runs the device instantiation of these inline functions.)
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "../grab_b200/csrc/swar.h"

using namespace gscan;

static uint32_t rnd_state = 12345;
static uint32_t rnd() { rnd_state = rnd_state * 1664525u + 1013904223u; return rnd_state; }

int main()
{
	// zero_bytes_superset: never misses a zero byte; exact for the lowest zero byte; zero result <=> no zero byte
	for (int it = 0; it < 2000000; it++) {
		uint32_t t = rnd();
		if (it & 1) t &= 0xff00ffffu; // force zero bytes often
		if (it & 2) t &= 0xffffff00u;
		const uint32_t f = zero_bytes_superset(t) & kHigh;
		bool any = false, low_found = false;
		for (int k = 0; k < 4; k++) {
			const bool z = ((t >> (8 * k)) & 0xff) == 0;
			const bool flag = (f >> (8 * k + 7)) & 1;
			any |= z;
			if (z && !flag) { printf("zero byte missed t=%08x\n", t); return 1; }
			if (!low_found && flag && !z) { printf("lowest flag not exact t=%08x\n", t); return 1; }
			if (z) low_found = true;
		}
		if ((f != 0) != any) { printf("any mismatch t=%08x\n", t); return 1; }
	}
	// pair_test: superset of per-byte masked equality
	for (int it = 0; it < 500000; it++) {
		const uint32_t w = rnd(), s = rnd();
		const uint8_t m0 = (it & 4) ? 0xff : 0xdf, v0 = (uint8_t)(rnd() & m0), m1 = (it & 8) ? 0xff : 0xf0, v1 = (uint8_t)(rnd() & m1);
		const uint32_t f = pair_test(w, s, m0 * kOnes, v0 * kOnes, m1 * kOnes, v1 * kOnes) & kHigh;
		for (int k = 0; k < 4; k++) {
			const bool want = (((w >> (8 * k)) & m0) == v0) && (((s >> (8 * k)) & m1) == v1);
			if (want && !((f >> (8 * k + 7)) & 1)) { printf("pair_test missed\n"); return 1; }
		}
	}
	// range7: exact for every lo <= hi in 0..127 and every byte value of x7
	for (int lo = 0; lo < 128; lo++)
		for (int hi = lo; hi < 128; hi++) {
			const uint32_t ge = (uint32_t)(0x80 - lo) * kOnes, gt = (uint32_t)(0x7f - hi) * kOnes;
			for (int b = 0; b < 128; b++) {
				const uint32_t x7 = (uint32_t)b * kOnes ^ ((uint32_t)((b * 7 + 3) & 127) << 8); // neighbouring byte differs
				const uint32_t r = range7(x7, ge, gt) & kHigh;
				const bool in0 = b >= lo && b <= hi;
				const int b1 = ((x7 >> 8) & 127);
				const bool in1 = b1 >= lo && b1 <= hi;
				if ((((r >> 7) & 1) != 0) != in0 || (((r >> 15) & 1) != 0) != in1) { printf("range7 wrong lo=%d hi=%d b=%d\n", lo, hi, b); return 1; }
			}
		}
	// pack_top_nibble: bit 7 of byte k -> bit 28 + k
	for (int m = 0; m < 16; m++) {
		uint32_t f = 0;
		for (int k = 0; k < 4; k++) if (m & (1 << k)) f |= 0x80u << (8 * k);
		if ((pack_top_nibble(f) >> 28) != (uint32_t)m) { printf("pack wrong %d\n", m); return 1; }
	}
	// runs_at_least against the definition
	for (int it = 0; it < 200000; it++) {
		uint32_t w = rnd() | rnd();
		if (it & 1) w |= rnd();
		for (uint32_t n = 1; n <= 17; n++) {
			const uint32_t got = runs_at_least(w, n) & 0xffffu;
			for (int i = 0; i < 16; i++) {
				bool all = true;
				for (uint32_t k = 0; k < n; k++) all = all && ((w >> (i + k)) & 1);
				if (all != (((got >> i) & 1) != 0)) { printf("runs_at_least wrong w=%08x n=%u i=%d\n", w, n, i); return 1; }
			}
		}
	}
	// the host-scheduled AND-with-shift steps the RUN kernel uses (engine.cu) give the same answer
	for (uint32_t nf = 1; nf <= 17; nf++) {
		uint32_t sh[5] = {0, 0, 0, 0, 0}, len = 1;
		int k = 0;
		while (len * 2 <= nf) { sh[k++] = len; len *= 2; }
		if (len < nf) sh[k++] = nf - len;
		for (int it = 0; it < 20000; it++) {
			uint32_t w = rnd() | rnd(), t = w;
			for (int i = 0; i < 5; i++) t &= t >> sh[i];
			if ((t & 0xffffu) != (runs_at_least(w, nf) & 0xffffu)) { printf("schedule wrong nf=%u\n", nf); return 1; }
		}
	}
	printf("swar ok\n");
	return 0;
}
