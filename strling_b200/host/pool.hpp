// A persistent pool of host threads for the extract pipeline.  Several pipeline stages (inflate, record walk, staging,
// replay) run at the same time on different batches; each of them hands the pool a job of `n` tasks and helps to work it
// off, so the machine's cores are shared by whatever stages have work instead of every stage spawning its own threads.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstddef>
#include <deque>
#include <exception>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <stdexcept>
#include <thread>
#include <vector>

namespace strling {

class Pool {
 public:
  explicit Pool(int threads) : n_(threads < 1 ? 1 : threads) {
    for (int i = 0; i + 1 < n_; i++) workers_.emplace_back([this]() { loop(); });
  }
  ~Pool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto &t : workers_) t.join();
  }
  Pool(const Pool &) = delete;
  Pool &operator=(const Pool &) = delete;
  int size() const { return n_; }

  // Runs f(task) for task = 0 .. n-1 on the pool and on the calling thread; returns when all of them are done.
  // An exception in a task is rethrown here (the first one).  Idle workers take tasks of the job with the highest
  // priority first: the later a stage sits in the pipeline the higher its priority, so that batches drain.
  void run(size_t n, const std::function<void(size_t)> &f, int priority = 0) {
    if (n == 0) return;
    if (n == 1 || n_ == 1) {   // inline, with the same contract: every task runs, the first exception is rethrown at the end
      std::string err;
      for (size_t i = 0; i < n; i++) {
        try {
          f(i);
        } catch (const std::exception &e) {
          if (err.empty()) err = e.what();
        }
      }
      if (!err.empty()) throw std::runtime_error(err);
      return;
    }
    auto job = std::make_shared<Job>();
    job->n = n;
    job->f = &f;
    job->priority = priority;
    {
      std::lock_guard<std::mutex> lk(mu_);
      auto it = jobs_.begin();
      while (it != jobs_.end() && (*it)->priority >= priority) ++it;
      jobs_.insert(it, job);
    }
    cv_.notify_all();
    work(*job);
    {
      std::unique_lock<std::mutex> lk(mu_);
      for (auto it = jobs_.begin(); it != jobs_.end(); ++it)
        if (it->get() == job.get()) { jobs_.erase(it); break; }
      job->cv.wait(lk, [&]() { return job->done == job->n; });
    }
    if (!job->error.empty()) throw std::runtime_error(job->error);
  }

  // f(begin, end, part) over `parts` contiguous ranges of [0, n)
  template <typename F>
  void ranges(size_t n, size_t parts, F f, int priority = 0) {
    if (parts < 1) parts = 1;
    if (parts > n) parts = n ? n : 1;
    const size_t per = (n + parts - 1) / parts;
    run(parts, [&](size_t p) {
      const size_t a = p * per, e = a + per < n ? a + per : n;
      if (a < e) f(a, e, p);
    }, priority);
  }

 private:
  struct Job {
    size_t n = 0;
    int priority = 0;
    const std::function<void(size_t)> *f = nullptr;
    std::atomic<size_t> next{0};
    size_t done = 0;  // under mu_
    std::string error;
    std::condition_variable cv;
  };

  // the thread that submitted the job works it off until no task is left to hand out
  void work(Job &j) {
    while (one_task(j)) {}
  }

  bool one_task(Job &j) {
    const size_t i = j.next.fetch_add(1, std::memory_order_relaxed);
    if (i >= j.n) return false;
    std::string err;
    try {
      (*j.f)(i);
    } catch (const std::exception &e) {
      err = e.what();
    }
    std::lock_guard<std::mutex> lk(mu_);
    j.done++;
    if (!err.empty() && j.error.empty()) j.error = err;
    if (j.done == j.n) j.cv.notify_all();
    return true;
  }

  // workers look at the job list again after every task, so a job of a later pipeline stage overtakes a long early one
  void loop() {
    while (true) {
      std::shared_ptr<Job> job;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&]() {
          if (stop_) return true;
          for (auto &j : jobs_)
            if (j->next.load(std::memory_order_relaxed) < j->n) return true;
          return false;
        });
        if (stop_) return;
        for (auto &j : jobs_)
          if (j->next.load(std::memory_order_relaxed) < j->n) { job = j; break; }
      }
      if (job) one_task(*job);
    }
  }

  int n_;
  std::vector<std::thread> workers_;
  std::mutex mu_;
  std::condition_variable cv_;
  std::deque<std::shared_ptr<Job>> jobs_;
  bool stop_ = false;
};

}  // namespace strling
