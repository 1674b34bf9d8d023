/*
 * selfplay_threads.cpp — CPU ORACLE (test infrastructure, NOT product code).
 *
 * Restates the reference's CPU self-play ARCHITECTURE (rust/src/self_play.rs:39-337) on top of the
 * C oracle's per-game functions, so that bench.py can time "the reference's way of doing it" on
 * the GPU box's host cores (the Rust crate itself cannot be built in this image: no cargo/rustc).
 *
 *   - one NN thread (self_play.rs:141-246): drains nn_queue into pending_games, groups the
 *     pending leaf positions by ModelID into sets of UNIQUE positions, evaluates at most
 *     max_nn_batch_size positions of the model with the most queued positions in one callback,
 *     fans results out to every pending game whose (model, leaf) was evaluated;
 *   - max(1, ncpu-1) MCTS threads (self_play.rs:252-337): on_received_policy, then either back
 *     to the NN queue, a temperature-sampled move, or to_result;
 *   - three queues (nn_queue, mcts_queue, done) and poison-pill shutdown (self_play.rs:325-332).
 */
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "c4a0_oracle.h"

namespace {

struct PosKey {
  uint64_t mask, value;
  bool operator==(const PosKey &o) const { return mask == o.mask && value == o.value; }
};
struct PosHash {
  size_t operator()(const PosKey &k) const {
    uint64_t x = k.mask * 0x9E3779B97F4A7C15ULL ^ (k.value + 0x7F4A7C15ULL + (k.mask << 6));
    x ^= x >> 29;
    return (size_t)(x * 0xBF58476D1CE4E5B9ULL);
  }
};

struct Game {
  c4o_game *g;
  size_t req_index;
};

struct EvalResult {
  float policy[7];
  float qp, qn;
};

struct Job {
  bool poison;
  Game game;
  EvalResult res;
};

/* crossbeam bounded channel stand-in: capacity is n_games everywhere in the reference
 * (self_play.rs:49-51) so sends never block; a plain locked deque has the same behaviour. */
template <typename T>
class Channel {
 public:
  void send(T v) {
    {
      std::lock_guard<std::mutex> l(m_);
      q_.push_back(std::move(v));
    }
    cv_.notify_one();
  }
  /* blocking receive; false when closed and empty */
  bool recv(T &out) {
    std::unique_lock<std::mutex> l(m_);
    cv_.wait(l, [&] { return !q_.empty() || closed_; });
    if (q_.empty()) return false;
    out = std::move(q_.front());
    q_.pop_front();
    return true;
  }
  bool try_recv(T &out) {
    std::lock_guard<std::mutex> l(m_);
    if (q_.empty()) return false;
    out = std::move(q_.front());
    q_.pop_front();
    return true;
  }
  void close() {
    {
      std::lock_guard<std::mutex> l(m_);
      closed_ = true;
    }
    cv_.notify_all();
  }

 private:
  std::mutex m_;
  std::condition_variable cv_;
  std::deque<T> q_;
  bool closed_ = false;
};

struct Shared {
  Channel<Game> nn_queue;
  Channel<Job> mcts_queue;
  std::atomic<size_t> n_games_remaining{0};
  std::atomic<int> mcts_threads_alive{0};
  std::atomic<int> error{0};
  std::atomic<uint64_t> sims{0}, nn_evals{0}, nn_batches{0}, moves{0}, samples{0};
  int max_nn_batch_size;
  uint64_t max_sims = 0; /* bench only: abandon the games once this many simulations ran (0 = play out) */
  uint64_t n_mcts_iterations;
  float c_exploration, c_ply_penalty;
  int n_mcts_threads;
  c4o_eval_fn eval;
  void *user;
  c4o_sample *out_samples;
  int *out_n;
};

/* self_play.rs:196-237 NNThread::loop_once (+ drain_queue :173-190, loop_until_close :241-245) */
void nn_thread(Shared *sh) {
  std::vector<Game> pending;
  bool chan_closed = false;
  std::vector<c4o_pos> pos;
  std::vector<float> pol, qp, qn;
  while (!chan_closed || !pending.empty()) {
    if (pending.empty()) {
      Game g;
      if (!sh->nn_queue.recv(g)) {
        chan_closed = true;
        continue;
      }
      pending.push_back(g);
    }
    {
      Game g;
      while (sh->nn_queue.try_recv(g)) pending.push_back(g);
    }
    if (pending.empty()) continue;

    std::map<uint64_t, std::unordered_set<PosKey, PosHash>> model_pos;
    for (const Game &game : pending) {
      c4o_pos p = c4o_game_leaf_pos(game.g);
      model_pos[c4o_game_leaf_model_id(game.g)].insert(PosKey{p.mask, p.value});
    }
    /* max_by_key over the BTreeMap: the last maximum wins */
    uint64_t model_id = 0;
    size_t best = 0;
    for (const auto &kv : model_pos)
      if (kv.second.size() >= best) {
        best = kv.second.size();
        model_id = kv.first;
      }
    pos.clear();
    for (const PosKey &k : model_pos[model_id]) {
      if ((int)pos.size() >= sh->max_nn_batch_size) break;
      pos.push_back(c4o_pos{k.mask, k.value});
    }
    size_t b = pos.size();
    pol.resize(b * 7);
    qp.resize(b);
    qn.resize(b);
    sh->eval(sh->user, model_id, (int)b, pos.data(), pol.data(), qp.data(), qn.data());
    sh->nn_evals += b;
    sh->nn_batches += 1;
    std::unordered_map<PosKey, size_t, PosHash> eval_map;
    eval_map.reserve(b * 2);
    for (size_t i = 0; i < b; i++) eval_map.emplace(PosKey{pos[i].mask, pos[i].value}, i);

    std::vector<Game> games;
    games.swap(pending);
    for (const Game &game : games) {
      c4o_pos p = c4o_game_leaf_pos(game.g);
      auto it = eval_map.find(PosKey{p.mask, p.value});
      if (c4o_game_leaf_model_id(game.g) != model_id || it == eval_map.end()) {
        pending.push_back(game);
        continue;
      }
      Job job;
      job.poison = false;
      job.game = game;
      std::memcpy(job.res.policy, &pol[it->second * 7], sizeof(float) * 7);
      job.res.qp = qp[it->second];
      job.res.qn = qn[it->second];
      sh->mcts_queue.send(job);
    }
  }
}

/* self_play.rs:268-337 MctsThread */
void mcts_thread(Shared *sh) {
  for (;;) {
    Job job;
    if (!sh->mcts_queue.recv(job)) break;
    if (job.poison) break;
    c4o_game *g = job.game.g;
    if (sh->max_sims && sh->sims.load() >= sh->max_sims) {
      /* time-boxed baseline run: the budget is spent, retire the game unfinished */
      c4o_game_free(g);
      if (sh->n_games_remaining.fetch_sub(1) == 1) {
        for (int i = 0; i < sh->n_mcts_threads - 1; i++) {
          Job pill;
          pill.poison = true;
          pill.game = Game{nullptr, 0};
          sh->mcts_queue.send(pill);
        }
        break;
      }
      continue;
    }
    sh->sims += 1;
    c4o_game_on_received_policy(g, job.res.policy, job.res.qp, job.res.qn, sh->c_exploration,
                                sh->c_ply_penalty);
    if (c4o_game_root_visit_count(g) < sh->n_mcts_iterations) {
      sh->nn_queue.send(job.game);
      continue;
    }
    c4o_pos root_pos = c4o_game_root_pos(g);
    if (c4o_terminal_state(root_pos) == C4O_NONE) {
      int ply = c4o_ply(root_pos);
      float temperature = ply < 4 ? 4.0f : (ply < 8 ? 2.0f : 1.0f);
      if (!c4o_game_make_random_move(g, sh->c_exploration, temperature)) sh->error = -2;
      sh->moves += 1;
      if (sh->error.load() == 0) {
        sh->nn_queue.send(job.game);
        continue;
      }
    }
    /* game over (or failed): to_result -> done */
    int n = c4o_game_to_result(g, sh->c_ply_penalty,
                               sh->out_samples + job.game.req_index * C4O_MAX_SAMPLES);
    sh->out_n[job.game.req_index] = n > 0 ? n : 0;
    if (n > 0) sh->samples += (uint64_t)n;
    c4o_game_free(g);
    if (sh->n_games_remaining.fetch_sub(1) == 1) {
      for (int i = 0; i < sh->n_mcts_threads - 1; i++) {
        Job pill;
        pill.poison = true;
        pill.game = Game{nullptr, 0};
        sh->mcts_queue.send(pill);
      }
      break;
    }
  }
  /* the last MCTS thread to leave closes nn_queue (all Senders dropped) */
  if (sh->mcts_threads_alive.fetch_sub(1) == 1) sh->nn_queue.close();
}

}  // namespace

extern "C" int c4o_self_play_threaded_budget(const c4o_metadata *reqs, size_t n_games,
                                             int max_nn_batch_size, uint64_t n_mcts_iterations,
                                             float c_exploration, float c_ply_penalty,
                                             c4o_eval_fn eval, void *user, int n_mcts_threads,
                                             uint64_t max_sims, c4o_sample *out_samples, int *out_n,
                                             c4o_stats *stats, uint64_t *nn_batches);

extern "C" int c4o_self_play_threaded(const c4o_metadata *reqs, size_t n_games,
                                      int max_nn_batch_size, uint64_t n_mcts_iterations,
                                      float c_exploration, float c_ply_penalty, c4o_eval_fn eval,
                                      void *user, int n_mcts_threads, c4o_sample *out_samples,
                                      int *out_n, c4o_stats *stats, uint64_t *nn_batches) {
  return c4o_self_play_threaded_budget(reqs, n_games, max_nn_batch_size, n_mcts_iterations,
                                       c_exploration, c_ply_penalty, eval, user, n_mcts_threads, 0,
                                       out_samples, out_n, stats, nn_batches);
}

/* Same, but stops after `max_sims` simulations (0 = never): a time-boxed sample of a workload whose
 * concurrency (n_games) matters for the NN batch size.  Unfinished games report out_n = 0; stats
 * count the moves made, each of which yields one training position at game end. */
extern "C" int c4o_self_play_threaded_budget(const c4o_metadata *reqs, size_t n_games,
                                             int max_nn_batch_size, uint64_t n_mcts_iterations,
                                             float c_exploration, float c_ply_penalty,
                                             c4o_eval_fn eval, void *user, int n_mcts_threads,
                                             uint64_t max_sims, c4o_sample *out_samples, int *out_n,
                                             c4o_stats *stats, uint64_t *nn_batches) {
  if (n_games == 0) {
    if (stats) std::memset(stats, 0, sizeof(*stats));
    return 0;
  }
  Shared sh;
  sh.max_nn_batch_size = max_nn_batch_size;
  sh.max_sims = max_sims;
  sh.n_mcts_iterations = n_mcts_iterations;
  sh.c_exploration = c_exploration;
  sh.c_ply_penalty = c_ply_penalty;
  if (n_mcts_threads <= 0) { /* self_play.rs:78: max(1, num_cpus - 1) */
    int hc = (int)std::thread::hardware_concurrency();
    n_mcts_threads = hc - 1 > 1 ? hc - 1 : 1;
  }
  sh.n_mcts_threads = n_mcts_threads;
  sh.eval = eval;
  sh.user = user;
  sh.out_samples = out_samples;
  sh.out_n = out_n;
  sh.n_games_remaining = n_games;
  sh.mcts_threads_alive = n_mcts_threads;
  c4o_pos empty = {0, 0};
  for (size_t i = 0; i < n_games; i++) {
    out_n[i] = 0;
    sh.nn_queue.send(Game{c4o_game_new(empty, reqs[i]), i});
  }
  std::thread nn(nn_thread, &sh);
  std::vector<std::thread> workers;
  for (int i = 0; i < n_mcts_threads; i++) workers.emplace_back(mcts_thread, &sh);
  for (auto &w : workers) w.join();
  nn.join();
  if (stats) {
    std::memset(stats, 0, sizeof(*stats));
    stats->sims = sh.sims;
    stats->nn_evals = sh.nn_evals;
    stats->moves = sh.moves;
    stats->samples = sh.samples;
  }
  if (nn_batches) *nn_batches = sh.nn_batches;
  return sh.error;
}
