// hash_check.cc -- CPU check of the perfect-hash tables the pattern compiler builds for the hashed scan engine
// (grab_b200/csrc/pattern.cc build_hash): the kernel's lookup, restated on the host, must find exactly the leading
// bytes of the alternatives -- every key, nothing else -- and the per-slot alternative lists must be the alternatives
// that start with the slot's key, in preference order.
#include <cstdio>
#include <cstdlib>
#include <set>
#include <string>
#include <vector>

#include "../grab_b200/csrc/pattern.h"

using namespace gscan;

static uint32_t umulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static uint64_t rng_state = 88172645463325252ull;
static uint32_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (uint32_t)(rng_state >> 11); }

// what HashEngine::probe computes (scan_kernels.cu): slot = umulhi(key * mul, slots); hit <=> table[slot] == key
static bool lookup(const Program &p, uint32_t key, uint32_t *slot)
{
	*slot = umulhi32(key * (p.hash_mul << (8 * (4 - p.hash_len))), p.hash_slots);
	return p.hash_table[*slot] == key;
}

static int check(const std::string &pat, bool expect_hash)
{
	Program p;
	std::string err;
	if (!compile_pattern(pat.data(), pat.size(), 0, p, err)) { printf("compile failed: %s\n", err.c_str()); return 1; }
	if (!p.use_hash) {
		if (expect_hash) { printf("pattern of %zu alternatives is not hashed\n", p.seqs.size()); return 1; }
		return 0;
	}
	const int L = p.hash_len;
	const uint32_t key_mask = L == 2 ? 0xffffu : 0xffffffu;
	if (L != 2 && L != 3) { printf("hash_len %d\n", L); return 1; }
	if (p.hash_slots == 0 || p.hash_slots > 8192 || p.hash_table.size() != p.hash_slots || p.slot_first.size() != p.hash_slots ||
	    p.slot_count.size() != p.hash_slots) { printf("table shape\n"); return 1; }
	// every leading-byte combination of every alternative is found, in the slot that lists the alternative
	std::set<uint32_t> keys;
	for (size_t si = 0; si < p.seqs.size(); si++) {
		std::vector<uint32_t> ks(1, 0);
		for (int i = 0; i < L; i++) {
			std::vector<uint32_t> next;
			for (uint32_t k : ks)
				for (unsigned b = 0; b < 256; b++)
					if (p.seqs[si][i].has(b)) next.push_back(k | (b << (8 * i)));
			ks.swap(next);
		}
		for (uint32_t k : ks) {
			keys.insert(k);
			uint32_t slot;
			if (!lookup(p, k, &slot)) { printf("key %06x of alternative %zu not found\n", k, si); return 1; }
			bool listed = false;
			uint32_t prev = 0;
			for (uint32_t j = 0; j < p.slot_count[slot]; j++) {
				const uint32_t s = p.slot_seqs[p.slot_first[slot] + j];
				if (j && s <= prev) { printf("slot list not in preference order\n"); return 1; }
				prev = s;
				listed = listed || s == si;
				// every listed alternative really starts with this key
				for (int i = 0; i < L; i++)
					if (!p.seqs[s][i].has((k >> (8 * i)) & 0xff)) { printf("slot lists a foreign alternative\n"); return 1; }
			}
			if (!listed) { printf("alternative %zu missing from its slot list\n", si); return 1; }
		}
	}
	// nothing else is found: all 2-byte values exhaustively, 3-byte values sampled + the neighbours of every key
	auto must_miss = [&](uint32_t k) -> bool {
		uint32_t slot;
		if (keys.count(k)) return true;
		if (lookup(p, k, &slot)) { printf("non-key %06x found\n", k); return false; }
		return true;
	};
	if (L == 2) {
		for (uint32_t k = 0; k < 65536; k++) if (!must_miss(k)) return 1;
	} else {
		for (int it = 0; it < 2000000; it++) if (!must_miss(rnd() & key_mask)) return 1;
		for (uint32_t k : keys)
			for (int bit = 0; bit < 24; bit++) if (!must_miss(k ^ (1u << bit))) return 1;
	}
	// empty slots are really empty and carry no list
	size_t filled = 0;
	for (uint32_t s = 0; s < p.hash_slots; s++) {
		if (p.hash_table[s] == 0xffffffffu) { if (p.slot_count[s]) { printf("empty slot with a list\n"); return 1; } }
		else { filled++; if (!keys.count(p.hash_table[s])) { printf("slot holds a non-key\n"); return 1; } }
	}
	if (filled != keys.size()) { printf("filled %zu != keys %zu\n", filled, keys.size()); return 1; }
	printf("ok: %zu alternatives, %zu keys of %d bytes, %u slots\n", p.seqs.size(), keys.size(), L, p.hash_slots);
	return 0;
}

static std::string random_set(int n, int lo, int hi, const char *alphabet, int na)
{
	std::set<std::string> lits;
	while ((int)lits.size() < n) {
		std::string s;
		const int len = lo + (int)(rnd() % (uint32_t)(hi - lo + 1));
		for (int i = 0; i < len; i++) s += alphabet[rnd() % (uint32_t)na];
		lits.insert(s);
	}
	std::string pat;
	for (auto &s : lits) { if (!pat.empty()) pat += '|'; pat += s; }
	return pat;
}

int main()
{
	int rc = 0;
	const char *lower = "abcdefghijklmnopqrstuvwxyz";
	rc |= check("alpha|bravo|charlie|delta|echo|foxtrot|golf|hotel", true);
	rc |= check("(?i)linus|torvalds|kernel|patch|merge", true);
	rc |= check("ab|ba|ca|cb|bc|ac|xy|zz|qq", true);                  // 2-byte keys
	rc |= check("a[ab]c|b[bc]a|c[ac]b|ab[ab]|ba[bc]|x[yz]z|q[a-d]q", true); // classes in the leading bytes
	for (int n : {9, 20, 50, 100, 150, 250}) rc |= check(random_set(n, 3, 6, lower, 26), true);
	for (int n : {10, 40}) rc |= check(random_set(n, 2, 4, lower, 26), true);   // shortest member has 2 bytes
	for (int n : {20, 50}) rc |= check(random_set(n, 3, 5, "ab", 2), false);    // tiny alphabet: few distinct keys (56 strings exist)
	if (rc == 0) printf("hash ok\n");
	return rc;
}
