// Host-side check of life_b200/host/rank_team.h (the one-thread-per-GPU team of the host programs), no GPU needed: every job runs on
// every rank exactly once, rank 0 on the calling thread, run() returns only after all ranks have finished, jobs do not overlap, and
// the destructor joins.  Built with -fsanitize=thread by tests/test_rank_team.py.
#include "rank_team.h"
#include <atomic>
#include <cstdio>

int main() {
	for (int n : {1, 2, 4, 8}) {
		RankTeam team;
		team.start(n);
		std::vector<long> hits((size_t)n, 0);
		std::atomic<int> inside{0};
		long plain = 0;                       // touched by rank 0 only, read by main between jobs: run() must order it
		const std::thread::id main_id = std::this_thread::get_id();
		bool rank0_on_caller = true;
		for (int job = 0; job < 2000; job++) {
			team.run([&](int r) {
				inside.fetch_add(1);
				hits[(size_t)r]++;            // each rank its own slot
				if (r == 0) { plain += job; if (std::this_thread::get_id() != main_id) rank0_on_caller = false; }
				inside.fetch_sub(1);
			});
			if (inside.load() != 0) { std::printf("FAIL: run() returned with a rank still inside (n=%d job=%d)\n", n, job); return 1; }
		}
		for (int r = 0; r < n; r++)
			if (hits[(size_t)r] != 2000) { std::printf("FAIL: rank %d ran %ld of 2000 jobs (n=%d)\n", r, hits[(size_t)r], n); return 1; }
		if (plain != 1999L * 2000 / 2 || !rank0_on_caller) { std::printf("FAIL: rank 0 bookkeeping (n=%d)\n", n); return 1; }
	}
	std::printf("OK\n");
	return 0;
}
