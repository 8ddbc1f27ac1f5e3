// rank_team.h — one host thread per GPU for the host programs of this directory (life_host.cpp: LIFE's own main() on N GPUs;
// life_run.cpp: the run-time front end).
#pragma once
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

// One host thread per GPU (the C ABI's rule: one thread per context; NCCL's rule: ranks of one process must not be driven from a
// single thread without group calls).  The reference's main() stays single threaded: each replaced body hands the same call to
// every rank's thread and waits for all of them, so to the reference's host code the N slabs look like one lattice.
class RankTeam {
public:
	void start(int n) {
		n_ = n;
		for (int r = 1; r < n; r++) workers_.emplace_back([this, r] { loop(r); });
	}
	~RankTeam() {
		if (abandon) { for (auto &t : workers_) t.detach(); return; }     // exit(99) from inside a rank's call: do not wait for the others
		{ std::lock_guard<std::mutex> g(m_); quit_ = true; gen_++; }
		cv_.notify_all();
		for (auto &t : workers_) t.join();
	}
	bool abandon = false;
	int size() const { return n_; }
	// fn(rank) on every rank concurrently (rank 0 on the calling thread); returns when all have finished
	void run(const std::function<void(int)> &fn) {
		if (n_ <= 1) { fn(0); return; }
		{ std::lock_guard<std::mutex> g(m_); job_ = &fn; pending_ = n_ - 1; gen_++; }
		cv_.notify_all();
		fn(0);
		std::unique_lock<std::mutex> g(m_);
		done_.wait(g, [this] { return pending_ == 0; });
		job_ = nullptr;
	}
private:
	void loop(int r) {
		unsigned long seen = 0;
		for (;;) {
			const std::function<void(int)> *job;
			{
				std::unique_lock<std::mutex> g(m_);
				cv_.wait(g, [&] { return gen_ != seen; });
				seen = gen_;
				if (quit_) return;
				job = job_;
			}
			(*job)(r);
			{ std::lock_guard<std::mutex> g(m_); pending_--; }
			done_.notify_one();
		}
	}
	int n_ = 1, pending_ = 0;
	unsigned long gen_ = 0;
	bool quit_ = false;
	const std::function<void(int)> *job_ = nullptr;
	std::mutex m_;
	std::condition_variable cv_, done_;
	std::vector<std::thread> workers_;
};

