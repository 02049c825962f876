// main.cc -- grab-b200: the reference's command line (/root/reference/src/main.cc:105-266) in front
// of the CUDA scan engine.  Same getopt string "Rrn:IOlsL", same config-map keys, same exit codes;
// additionally accepts -2 -H (aliases: same PCRE match semantics) and -S (literal pattern) from the
// reference's README.md:16-31.  GPU knobs are environment variables so the short-flag surface
// stays identical: GRAB_B200_DEVICE=<n> (first GPU), GRAB_B200_NDEV=<k> (GPUs to spread over: threads under -n,
// batches of windows otherwise -- also the windows of ONE huge file), GRAB_B200_LANES=<k> (scan lanes per GPU),
// GRAB_B200_BATCH_BYTES=<n> (bytes of windows per engine call, default 256 MiB), GRAB_B200_LENIENT=1 (do not
// reproduce quirk Q2).
#include <ftw.h>
#include <pthread.h>
#include <sched.h>
#include <unistd.h>

#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <new>
#include <string>
#include <vector>

#include "filegrep.h"

using namespace std;
using grab_b200::FileGrep;

struct thread_arg {
	int idx, nthreads;
	FileGrep *grep;
};

static vector<string> files;
static vector<struct stat> stats;

static int thread_walk(const char *path, const struct stat *st, int typeflag, struct FTW *)
{
	if (typeflag == FTW_F && S_ISREG(st->st_mode)) { // main.cc:74-83
		files.push_back(path);
		stats.push_back(*st);
	}
	return 0;
}

static void *find_iterative(void *vp)
{
	thread_arg *ta = static_cast<thread_arg *>(vp);
	int vsize = (int)files.size();
	for (int i = ta->idx; i < vsize; i += ta->nthreads) // static round-robin over files, main.cc:94
		ta->grep->find(files[i].c_str(), &stats[i], FTW_F);
	ta->grep->flush();
	return nullptr;
}

static void usage(const string &p)
{
	cout << "Usage: " << p << " [-rR] [-I] [-O] [-L] [-l] [-s] [-n <cores>] <regex> <path>\n";
	exit(1);
}

int main(int argc, char **argv)
{
	int c = 0;
	map<string, size_t> config;
	size_t chunk_size = (size_t)1 << 30;

	while ((c = getopt(argc, argv, "Rrn:IOlsL2HS")) != -1) {
		switch (c) {
		case 'r': case 'R': config["recursive"] = 1; break;
		case 's': config["single"] = 1; break;
		case 'O': config["offsets"] = 1; break;
		case 'l': config["noline"] = 1; break;
		case 'L':
			config["low_mem"] = 1;
			chunk_size >>= 1;
			if (chunk_size < ((size_t)1 << 25)) chunk_size = (size_t)1 << 25;
			break;
		case 'I':
			if (isatty(1)) config["color"] = 1;
			break;
		case 'n': config["cores"] = (size_t)atoi(optarg); break;
		case '2': case 'H': break; // engine selectors of the greppin branch: one engine here
		case 'S': config["literal"] = 1; break;
		default: usage(argv[0]);
		}
	}
	config["chunk_size"] = chunk_size;
	if (getenv("GRAB_B200_LENIENT")) config["lenient"] = 1;
	int device = getenv("GRAB_B200_DEVICE") ? atoi(getenv("GRAB_B200_DEVICE")) : 0;
	int ndev = getenv("GRAB_B200_NDEV") ? atoi(getenv("GRAB_B200_NDEV")) : 1;
	if (ndev < 1) ndev = 1;
	if (getenv("GRAB_B200_BATCH_BYTES") && atoll(getenv("GRAB_B200_BATCH_BYTES")) > 0) config["batch_bytes"] = (size_t)atoll(getenv("GRAB_B200_BATCH_BYTES"));
	if (getenv("GRAB_B200_LANES") && atoi(getenv("GRAB_B200_LANES")) > 0) config["lanes"] = (size_t)atoi(getenv("GRAB_B200_LANES"));

	if (argc < optind + 2) usage(argv[0]);
	string regex = argv[optind++];
	string path = argv[optind++];

	int cores = (int)config["cores"];
	if (cores > 1) {
		if (config.count("recursive") == 0) {
			cerr << "Multicore support only for recursive grabs.\n";
			return -1;
		}
		chunk_size >>= 2; // main.cc:172-173
		// every thread stages its own batches: keep the engine's helper lanes from oversubscribing the host
		if (!getenv("GSCAN_STAGE_THREADS")) setenv("GSCAN_STAGE_THREADS", cores >= 8 ? "1" : "2", 1);
		config["chunk_size"] = chunk_size;
		files.reserve(1 << 20);
		stats.reserve(1 << 20);
		nftw(path.c_str(), thread_walk, 1024, FTW_PHYS);

		thread_arg *ta = new (nothrow) thread_arg[cores];
		pthread_t *tids = new (nothrow) pthread_t[cores];
		if (!ta || !tids) { cerr << "Out of memory.\n"; return -1; }
		for (int i = 0; i < cores; ++i) {
			// one private engine per thread (main.cc:195-199); threads spread over the box's GPUs
			map<string, size_t> cfg = config;
			cfg["device"] = (size_t)(device + i % ndev);
			FileGrep *tgrep = new (nothrow) FileGrep;
			tgrep->config(cfg);
			tgrep->prepare(regex); // return value ignored, as in main.cc:198
			tgrep->recurse();
			ta[i].grep = tgrep;
			ta[i].idx = i;
			ta[i].nthreads = cores;
			int r = 0;
			if ((r = pthread_create(tids + i, nullptr, find_iterative, ta + i)) != 0) {
				cerr << "pthread_create: " << strerror(r) << endl;
				exit(-1);
			}
			cpu_set_t cpuset;
			CPU_ZERO(&cpuset);
			CPU_SET(i, &cpuset);
			if ((r = pthread_setaffinity_np(tids[i], sizeof(cpuset), &cpuset)) != 0) {
				cerr << "pthread_setaffinity_np:" << strerror(r) << " (more threads than cores?)" << endl;
				exit(-1);
			}
		}
		for (int i = 0; i < cores; ++i) {
			pthread_join(tids[i], nullptr);
			delete ta[i].grep;
		}
		delete[] ta;
		delete[] tids;
		exit(0);
	}

	config["device"] = (size_t)device;
	config["ndev"] = (size_t)ndev; // batches of windows go round the GPUs; stdout order does not depend on it
	FileGrep *grep = new (nothrow) FileGrep;
	if (!grep) { cerr << "Out of memory.\n"; return -1; }
	grep->config(config);
	if (grep->prepare(regex) < 0) {
		cerr << grep->why() << endl;
		return -1;
	}
	if (config.count("recursive") > 0) {
		if (grep->find_recursive(path) < 0) {
			cerr << grep->why() << endl;
			return -1;
		}
	} else {
		if (argc - optind > 0) grep->show_path(1);
		for (;;) {
			if (grep->find(path) < 0) {
				cerr << grep->why() << endl;
				return -1;
			}
			if (argc > optind) path = argv[optind++];
			else break;
		}
		if (grep->flush() < 0) {
			cerr << grep->why() << endl;
			return -1;
		}
	}
	delete grep;
	return 0;
}
