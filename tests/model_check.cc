// model_check.cc -- CPU check of what the pattern compiler (grab_b200/csrc/pattern.cc) hands to the device, against
// the oracle: the compiled Program is interpreted on the host exactly as the resolve kernels interpret it -- FIXED:
// first sequence in preference order at the leftmost position; RUN: greedy class run of at least run_min; general
// patterns: the device VM, whose source is extracted verbatim from resolve_kernels.cu by tests/test_model.py
// (vm_snippet.inc) and compiled here for the host -- inside the reference's loop (moving search start, strict '<'
// guard, grab.cc:175-213).  Also checks that the candidate filter of general patterns (leading-byte sequences / run
// starts) never rejects a position where the VM matches.
// Usage: model_check PATTERN_FILE   (one pattern per line)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "../grab_b200/csrc/pattern.h"
extern "C" {
#include "../oracle/grab_oracle.h"
}

namespace gscan {
struct ResolveArgs { const uint32_t *vm_code; const uint32_t *vm_sets; };
#define __device__
#define __forceinline__ inline
#include "vm_snippet.inc"
} // namespace gscan

using namespace gscan;

static uint64_t rng_state = 0x1234567ull;
static uint32_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (uint32_t)(rng_state >> 9); }

struct M { uint64_t pos; uint32_t len; };

static bool seq_at(const Sequence &q, const uint8_t *s, size_t len, size_t pos)
{
	if (pos + q.size() > len) return false;
	for (size_t i = 0; i < q.size(); i++) if (!q[i].has(s[pos + i])) return false;
	return true;
}

// 0 ok, -1 VM limit, -2 filter soundness violated
static int emulate(const Program &p, const uint8_t *s, size_t len, std::vector<M> &out, std::string &why)
{
	ResolveArgs R{p.vm_code.data(), p.vm_sets.data()};
	size_t start = 0;
	while (start + (size_t)p.minlen < len) { // grab.cc:175
		bool found = false;
		size_t pos = start, mend = 0;
		for (; pos < len; pos++) {
			if (p.use_vm) {
				uint32_t e = 0;
				const int rc = vm_exec(R, s + start, (uint32_t)(len - start), (uint32_t)(pos - start), &e);
				if (rc < 0) return -1;
				if (rc == 1) { found = true; mend = start + e; }
			} else if (p.kind == ENGINE_RUN) {
				size_t k = 0;
				while (pos + k < len && p.run_class.has(s[pos + k])) k++;
				if (k >= (size_t)p.run_min) { found = true; mend = pos + k; }
			} else {
				for (auto &q : p.seqs)
					if (seq_at(q, s, len, pos)) { found = true; mend = pos + q.size(); break; }
			}
			if (found) break;
		}
		if (!found) break;
		if (p.use_vm) { // the scan kernel must have offered this position to the VM walk
			bool offered = false;
			if (p.vm_runstart) offered = p.run_class.has(s[pos]) && (pos == start || !p.run_class.has(s[pos - 1]));
			else for (auto &q : p.seqs) offered = offered || seq_at(q, s, len, pos);
			if (!offered) { why = "candidate filter rejects a matching position"; return -2; }
		}
		if (mend <= pos) { why = "empty match"; return -2; }
		out.push_back(M{pos, (uint32_t)(mend - pos)});
		start = mend;
	}
	return 0;
}

int main(int argc, char **argv)
{
	if (argc < 2) return 2;
	std::ifstream f(argv[1]);
	std::string pat;
	std::vector<std::vector<uint8_t>> subjects;
	const char *alphas[] = {"abc", "abcx \n", "ab", "abc abc\n\n", "aabbcc_x1 \t\n", "ab\n"};
	const int lens[] = {1, 2, 3, 4, 5, 7, 9, 16, 17, 33, 64, 130, 257};
	for (int k = 0; k < 78; k++) {
		const char *al = alphas[k % 6];
		const size_t na = strlen(al), ln = (size_t)lens[k % 13];
		std::vector<uint8_t> b(ln);
		for (auto &c : b) c = (uint8_t)al[rnd() % na];
		subjects.push_back(b);
	}
	int n_pat = 0, n_served = 0, n_vm = 0, n_cmp = 0, n_limit = 0, bad = 0;
	while (std::getline(f, pat)) {
		if (pat.empty()) continue;
		n_pat++;
		Program p;
		std::string err;
		if (!compile_pattern(pat.data(), pat.size(), 0, p, err)) continue; // rejected loudly: nothing to check
		char oerr[200];
		go_regex *re = go_compile(pat.data(), pat.size(), 0, oerr, sizeof oerr);
		if (!re) continue;
		if (go_nullable(re)) { printf("MISMATCH %s: compiled but the oracle says it can match the empty string\n", pat.c_str()); bad++; go_free(re); continue; }
		if (go_minlen(re) != p.minlen) { printf("MISMATCH %s: minlen %d, oracle %d\n", pat.c_str(), p.minlen, go_minlen(re)); bad++; }
		n_served++;
		n_vm += p.use_vm;
		for (auto &sb : subjects) {
			go_matches want = {0, 0, 0};
			if (go_scan_window(re, sb.data(), sb.size(), 0, 0, GO_MODE_ALL, 0, &want) != 0) { go_matches_free(&want); n_limit++; continue; }
			std::vector<M> got;
			std::string why;
			const int rc = emulate(p, sb.data(), sb.size(), got, why);
			if (rc == -1) { n_limit++; go_matches_free(&want); continue; }
			bool same = rc == 0 && got.size() == want.n;
			for (size_t i = 0; same && i < got.size(); i++) same = got[i].pos == want.v[i].start && got[i].len == want.v[i].len;
			n_cmp++;
			if (!same) {
				bad++;
				printf("MISMATCH %s (engine %d vm %d runstart %d) on \"", pat.c_str(), (int)p.kind, (int)p.use_vm, (int)p.vm_runstart);
				for (uint8_t c : sb) printf(c == '\n' ? "\\n" : c == '\t' ? "\\t" : "%c", c);
				printf("\": %s got %zu want %zu\n", why.c_str(), got.size(), want.n);
				if (bad > 20) { go_matches_free(&want); go_free(re); return 1; }
			}
			go_matches_free(&want);
		}
		go_free(re);
	}
	printf("patterns %d, served %d (%d through the VM), comparisons %d, limit skips %d, mismatches %d\n", n_pat, n_served, n_vm, n_cmp, n_limit, bad);
	if (bad == 0) printf("model ok\n");
	return bad ? 1 : 0;
}
