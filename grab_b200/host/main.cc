// main.cc -- grab-b200, the command line of the B200 scan engine.
//
// Written against the BEHAVIOUR of the reference's front end (flag letters, messages and exit codes recorded in
// tests/golden/kat.json from the unmodified binary; semantics in /root/reference/README.md:16-31 and SURVEY.md 8(b)
// "CLI surface to keep"), not against its source: options are table driven, the -n crew is a small pool class, and
// the directory walk that feeds the crew is parallel (the reference walks serially before it starts its threads,
// README.md:137-139 -- on a GPU box that serial walk is most of the wall time of a recursive grab).
//
//   grab-b200 [-rR] [-I] [-O] [-L] [-l] [-s] [-n <cores>] [-S] [-2] [-H] <regex> <path> [<path> ...]
//
// GPU knobs are environment variables, so the short-flag surface stays the reference's:
//   GRAB_B200_DEVICE=<n>       first GPU
//   GRAB_B200_NDEV=<k>         GPUs to spread over (threads under -n, batches of windows otherwise)
//   GRAB_B200_LANES=<k>        scan lanes per GPU
//   GRAB_B200_BATCH_BYTES=<n>  bytes of windows per engine call (default 256 MiB)
//   GRAB_B200_LENIENT=1        do not reproduce quirk Q2 (capturing groups print nothing)
#include <dirent.h>
#include <fcntl.h>
#include <pthread.h>
#include <sched.h>
#include <sys/resource.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <iostream>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "filegrep.h"

namespace {

using grab_b200::FileGrep;
typedef std::map<std::string, size_t> Settings;

constexpr int kFailure = -1;                     // what the reference returns from main: exit status 255
constexpr size_t kDefaultChunk = (size_t)1 << 30; // grab.h:48
constexpr size_t kSmallestChunk = (size_t)1 << 25;

struct Invocation {
	Settings settings;
	size_t chunk = kDefaultChunk;
	int workers = 0;
	std::string regex;
	std::vector<std::string> paths;
	int first_gpu = 0, gpus = 1;
};

[[noreturn]] void print_usage_and_exit(const char *self)
{
	std::cout << "Usage: " << self << " [-rR] [-I] [-O] [-L] [-l] [-s] [-n <cores>] <regex> <path>\n";
	exit(1);
}

// ---- options: one row per flag letter ------------------------------------------------------------------------
struct FlagRow {
	char letter;
	const char *key;      // settings key switched on by the flag (nullptr: see `special`)
	void (*special)(Invocation &, const char *arg);
};

void halve_chunk(Invocation &inv, const char *)
{
	inv.settings["low_mem"] = 1;
	inv.chunk = inv.chunk / 2 < kSmallestChunk ? kSmallestChunk : inv.chunk / 2; // every -L halves, never below 32 MiB
}
void colour_if_terminal(Invocation &inv, const char *) { if (isatty(STDOUT_FILENO)) inv.settings["color"] = 1; }
void set_workers(Invocation &inv, const char *arg) { inv.workers = atoi(arg); inv.settings["cores"] = (size_t)inv.workers; }
void same_engine(Invocation &, const char *) {} // -2 / -H pick engines of the greppin branch: one engine here

const FlagRow kFlags[] = {
	{'r', "recursive", nullptr}, {'R', "recursive", nullptr}, {'s', "single", nullptr}, {'O', "offsets", nullptr},
	{'l', "noline", nullptr},    {'S', "literal", nullptr},   {'L', nullptr, halve_chunk}, {'I', nullptr, colour_if_terminal},
	{'n', nullptr, set_workers}, {'2', nullptr, same_engine}, {'H', nullptr, same_engine},
};

size_t env_number(const char *name, size_t fallback)
{
	const char *v = getenv(name);
	if (!v || !*v) return fallback;
	const long long n = atoll(v);
	return n > 0 ? (size_t)n : fallback;
}

Invocation parse(int argc, char **argv)
{
	Invocation inv;
	for (int c; (c = getopt(argc, argv, "Rrn:IOlsL2HS")) != -1;) {
		const FlagRow *row = nullptr;
		for (const FlagRow &f : kFlags) if (f.letter == c) row = &f;
		if (!row) print_usage_and_exit(argv[0]);
		if (row->key) inv.settings[row->key] = 1;
		else row->special(inv, optarg);
	}
	if (argc - optind < 2) print_usage_and_exit(argv[0]);
	inv.regex = argv[optind++];
	while (optind < argc) inv.paths.push_back(argv[optind++]);
	if (getenv("GRAB_B200_LENIENT")) inv.settings["lenient"] = 1;
	inv.first_gpu = (int)env_number("GRAB_B200_DEVICE", 0);
	inv.gpus = (int)env_number("GRAB_B200_NDEV", 1);
	if (size_t b = env_number("GRAB_B200_BATCH_BYTES", 0)) inv.settings["batch_bytes"] = b;
	if (size_t l = env_number("GRAB_B200_LANES", 0)) inv.settings["lanes"] = l;
	return inv;
}

// ---- parallel directory walk -----------------------------------------------------------------------------------
// Same visit rule as the reference's nftw(..., FTW_PHYS): regular files only, symbolic links are never followed,
// unreadable directories are skipped without a word.  Directories are a shared work list; every walker keeps the
// files it finds in its own list (no lock on the hot path), and each file carries the stat the scan needs.
struct FoundFile {
	std::string path;
	struct stat st;
};

class TreeWalk {
public:
	explicit TreeWalk(int walkers) : d_found((size_t)(walkers < 1 ? 1 : walkers)) {}

	void run(const std::string &root)
	{
		struct stat st;
		if (lstat(root.c_str(), &st) != 0) return; // the reference ignores nftw's verdict in this mode
		if (S_ISREG(st.st_mode)) { d_found[0].push_back(FoundFile{root, st}); return; }
		if (!S_ISDIR(st.st_mode)) return;
		d_dirs.push_back(root);
		std::vector<pthread_t> tids(d_found.size());
		std::vector<Arg> args(d_found.size());
		size_t started = 0;
		for (; started < tids.size(); started++) {
			args[started] = Arg{this, started};
			if (pthread_create(&tids[started], nullptr, &TreeWalk::entry, &args[started]) != 0) break;
		}
		if (started == 0) walker(0); // no helper could be started: walk on this thread
		for (size_t i = 0; i < started; i++) pthread_join(tids[i], nullptr);
	}

	// all files, walker by walker (the order is as arbitrary as readdir's; output parity is defined on sorted lines)
	std::vector<FoundFile> take()
	{
		std::vector<FoundFile> all;
		size_t n = 0;
		for (auto &v : d_found) n += v.size();
		all.reserve(n);
		for (auto &v : d_found) { for (auto &f : v) all.push_back(std::move(f)); v.clear(); }
		return all;
	}

private:
	struct Arg { TreeWalk *self; size_t id; };
	static void *entry(void *p) { Arg *a = static_cast<Arg *>(p); a->self->walker(a->id); return nullptr; }

	bool next_dir(std::string &dir)
	{
		std::unique_lock<std::mutex> lk(d_mu);
		for (;;) {
			if (!d_dirs.empty()) { dir = std::move(d_dirs.back()); d_dirs.pop_back(); d_busy++; return true; }
			if (d_busy == 0) { d_cv.notify_all(); return false; } // nothing queued, nobody can add: done
			d_cv.wait(lk);
		}
	}

	void walker(size_t id)
	{
		std::string dir;
		std::vector<std::string> sub;
		while (next_dir(dir)) {
			sub.clear();
			if (DIR *d = opendir(dir.c_str())) {
				const int dfd = dirfd(d);
				while (struct dirent *e = readdir(d)) {
					const char *n = e->d_name;
					if (n[0] == '.' && (n[1] == 0 || (n[1] == '.' && n[2] == 0))) continue;
					if (e->d_type == DT_DIR) { sub.push_back(dir + "/" + n); continue; }
					if (e->d_type != DT_REG && e->d_type != DT_UNKNOWN) continue; // links, devices, sockets: never scanned
					struct stat st;
					if (fstatat(dfd, n, &st, AT_SYMLINK_NOFOLLOW) != 0) continue;
					if (S_ISDIR(st.st_mode)) sub.push_back(dir + "/" + n);
					else if (S_ISREG(st.st_mode)) d_found[id].push_back(FoundFile{dir + "/" + n, st});
				}
				closedir(d);
			}
			std::lock_guard<std::mutex> g(d_mu);
			for (auto &s : sub) d_dirs.push_back(std::move(s));
			d_busy--;
			d_cv.notify_all();
		}
	}

	std::mutex d_mu;
	std::condition_variable d_cv;
	std::vector<std::string> d_dirs;
	int d_busy = 0;
	std::vector<std::vector<FoundFile>> d_found;
};

// ---- the -n crew: N scanners, scanner i pinned to CPU i, each with a private engine ---------------------------------
class Crew {
public:
	Crew(const Invocation &inv, const std::vector<FoundFile> &files) : d_inv(inv), d_files(files) {}

	// returns only when every scanner is done; thread-start problems end the process like the reference does
	void run()
	{
		const int n = d_inv.workers;
		std::vector<Member> crew((size_t)n);
		for (int i = 0; i < n; i++) {
			Member &m = crew[(size_t)i];
			m.owner = this;
			m.rank = i;
			Settings s = d_inv.settings;
			s["device"] = (size_t)(d_inv.first_gpu + i % d_inv.gpus); // scanners spread over the box's GPUs
			m.grep.reset(new FileGrep);
			m.grep->config(s);
			m.grep->prepare(d_inv.regex); // a bad pattern is not reported in this mode (the scanners then find nothing)
			m.grep->recurse();
			if (int r = pthread_create(&m.tid, nullptr, &Crew::entry, &m)) {
				std::cerr << "pthread_create: " << strerror(r) << std::endl;
				exit(kFailure);
			}
			cpu_set_t one;
			CPU_ZERO(&one);
			CPU_SET(i, &one);
			if (int r = pthread_setaffinity_np(m.tid, sizeof(one), &one)) {
				std::cerr << "pthread_setaffinity_np:" << strerror(r) << " (more threads than cores?)" << std::endl;
				exit(kFailure);
			}
		}
		for (Member &m : crew) {
			pthread_join(m.tid, nullptr);
			m.grep->flush();          // prints what is still queued
			(void)m.grep.release();   // lanes, engine contexts and pinned buffers go with the process (see main)
		}
	}

private:
	struct Member {
		Crew *owner = nullptr;
		int rank = 0;
		pthread_t tid{};
		std::unique_ptr<FileGrep> grep;
	};
	static void *entry(void *p)
	{
		Member *m = static_cast<Member *>(p);
		const std::vector<FoundFile> &files = m->owner->d_files;
		const size_t stride = (size_t)m->owner->d_inv.workers;
		for (size_t i = (size_t)m->rank; i < files.size(); i += stride) // file i belongs to scanner i mod N
			m->grep->find(files[i].path.c_str(), &files[i].st, 0);
		m->grep->flush();
		return nullptr;
	}
	const Invocation &d_inv;
	const std::vector<FoundFile> &d_files;
};

int run_crew(Invocation &inv)
{
	if (inv.settings.count("recursive") == 0) {
		std::cerr << "Multicore support only for recursive grabs.\n";
		return kFailure;
	}
	inv.settings["chunk_size"] = inv.chunk / 4; // N scanners hold N windows: a quarter of the chunk each
	// every scanner stages its own batches: keep the engine's helper lanes from oversubscribing the host
	if (!getenv("GSCAN_STAGE_THREADS")) setenv("GSCAN_STAGE_THREADS", inv.workers >= 8 ? "1" : "2", 1);
	TreeWalk walk(inv.workers < 32 ? inv.workers : 32);
	walk.run(inv.paths[0]);
	const std::vector<FoundFile> files = walk.take();
	Crew(inv, files).run();
	return 0;
}

int run_single(Invocation &inv)
{
	inv.settings["device"] = (size_t)inv.first_gpu;
	inv.settings["ndev"] = (size_t)inv.gpus; // batches of windows go round the GPUs; stdout order does not depend on it
	std::unique_ptr<FileGrep> grep(new (std::nothrow) FileGrep);
	if (!grep) { std::cerr << "Out of memory.\n"; return kFailure; }
	grep->config(inv.settings);
	int status = 0;
	if (grep->prepare(inv.regex) < 0) {
		std::cerr << grep->why() << std::endl;
		return kFailure;
	}
	if (inv.settings.count("recursive")) {
		if (grep->find_recursive(inv.paths[0]) < 0) { std::cerr << grep->why() << std::endl; status = kFailure; }
		(void)grep.release(); // everything is printed (find_recursive flushes): lanes, engine contexts and pinned buffers go with the process
		return status;
	}
	if (inv.paths.size() > 1) grep->show_path(true);
	for (const std::string &p : inv.paths) {
		if (grep->find(p) < 0) {
			// what earlier paths queued is printed before the complaint, as the reference (which scans path by path) does
			grep->flush();
			std::cerr << grep->why() << std::endl;
			return kFailure;
		}
	}
	if (grep->flush() < 0) { std::cerr << grep->why() << std::endl; status = kFailure; }
	(void)grep.release();
	return status;
}


// Descriptor feed (no line output: windows are read by the engine's staging threads, not mapped): every queued window
// holds a descriptor, so the soft limit is raised to the hard one and shared out over the scanners.
// GRAB_B200_FEED=mmap keeps the reference's mappings.
size_t descriptor_budget(int scanners)
{
	const char *feed = getenv("GRAB_B200_FEED");
	if (feed && strcmp(feed, "mmap") == 0) return 0;
	struct rlimit rl;
	if (getrlimit(RLIMIT_NOFILE, &rl) != 0) return 0;
	const rlim_t want = rl.rlim_max == RLIM_INFINITY || rl.rlim_max > 65536 ? 65536 : rl.rlim_max;
	if (rl.rlim_cur < want) {
		struct rlimit up = rl;
		up.rlim_cur = want;
		if (setrlimit(RLIMIT_NOFILE, &up) == 0) rl.rlim_cur = want;
	}
	if (rl.rlim_cur == RLIM_INFINITY) rl.rlim_cur = 65536;
	return rl.rlim_cur > 128 ? (size_t)(rl.rlim_cur - 128) / (size_t)(scanners < 1 ? 1 : scanners) : 0;
}

} // namespace

int main(int argc, char **argv)
{
	Invocation inv = parse(argc, argv);
	inv.settings["chunk_size"] = inv.chunk;
	inv.settings["fd_budget"] = descriptor_budget(inv.workers);
	const int status = inv.workers > 1 ? run_crew(inv) : run_single(inv);
	// Everything is printed: leave without closing the engine contexts (two dozen cudaFreeHost / cudaFree calls) and without
	// the CUDA runtime's exit handlers -- 0.2 s of orderly teardown (tools/feed_bench.py: last print at 599 ms, process gone at
	// 816 ms) that the kernel's own cleanup at process exit makes redundant
	std::cout.flush();
	std::cerr.flush();
	fflush(nullptr);
	_exit(status);
}
