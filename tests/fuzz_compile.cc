// fuzz_compile.cc -- robustness of the pattern compiler (grab_b200/csrc/pattern.cc) and of the oracle's parser on
// metacharacter soup: no crash, no hang, and wherever both accept a pattern they agree on MINLENGTH and on whether it
// can match the empty string (the compiler must reject those, Q4).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "../grab_b200/csrc/pattern.h"
extern "C" {
#include "../oracle/grab_oracle.h"
}

static uint64_t st = 0x9E3779B97F4A7C15ull;
static uint32_t rnd() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (uint32_t)(st >> 13); }

int main(int argc, char **argv)
{
	const int n = argc > 1 ? atoi(argv[1]) : 200000;
	const char *toks[] = {"a", "b", "c", "x", ".", "*", "+", "?", "|", "(", ")", "(?:", "(?i)", "(?x)", "(?U)", "(?#", "[", "]", "[^", "-", "{", "}", "{2}", "{1,3}",
	                      "{2,}", ",", "\\", "\\w", "\\d", "\\s", "\\b", "\\B", "^", "$", "\\A", "\\z", "\\x41", "\\x{", "\\Q", "\\E", " ", "#", "\n", "1", "9",
	                      "[a-c]", "[[:alpha:]]", "(?s)", "(?m)", "(?-i)", "*?", "+?", "??", "*+", "++", "\\1", "(?=", "(?<", "\xff", "\\", "[:", ":]"};
	const int nt = (int)(sizeof toks / sizeof toks[0]);
	int accepted = 0, both = 0, bad = 0;
	double worst = 0;
	std::string worst_pat;
	for (int it = 0; it < n; it++) {
		std::string pat;
		const int len = 1 + (int)(rnd() % 12);
		for (int i = 0; i < len; i++) pat += toks[rnd() % (uint32_t)nt];
		const auto t0 = std::chrono::steady_clock::now();
		gscan::Program p;
		std::string err;
		const bool ok = gscan::compile_pattern(pat.data(), pat.size(), 0, p, err);
		const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
		if (ms > worst) { worst = ms; worst_pat = pat; }
		if (!ok && err.empty()) { printf("rejected without a message: %s\n", pat.c_str()); bad++; }
		char oerr[200];
		go_regex *re = go_compile(pat.data(), pat.size(), 0, oerr, sizeof oerr);
		if (ok) accepted++;
		if (ok && re) {
			both++;
			if (go_nullable(re)) { printf("compiler accepts an empty-matchable pattern: %s\n", pat.c_str()); bad++; }
			else if (go_minlen(re) != p.minlen) { printf("minlen %d vs oracle %d: %s\n", p.minlen, go_minlen(re), pat.c_str()); bad++; }
		}
		if (re) go_free(re);
		if (bad > 20) break;
	}
	printf("%d patterns, %d accepted, %d accepted by both, slowest compile %.1f ms, problems %d\n", n, accepted, both, worst, bad);
	if (worst > 2000) { printf("compile time out of bounds for: %s\n", worst_pat.c_str()); bad++; }
	if (!bad) printf("fuzz ok\n");
	return bad ? 1 : 0;
}
