// C++ host test of the sharded path behind the C ABI (include/apgk.h "a GROUP of ranks"), without Python or torch.
//
//   test_group local N         one process, N contexts on device 0 (apgk_group_local): runs on a one-GPU box
//   test_group procs N         N processes, one GPU each (fork; rank 0's apgk_group_unique_id travels through pipes,
//                              every rank calls apgk_group_join: NCCL for the small collectives, CUDA IPC peer memory
//                              for the exchange).  Needs N GPUs; exits 77 ("skipped") when there are fewer.
//
// Each rank generates its slice of one synthetic read set (apgk_synth_reads), the group counts it; the result must
// equal ONE context counting all the reads (apgk_finish): totals, the whole spectrum, and every shard's table must be
// found record for record in the single-GPU table.  A second step with fresh, larger read sets follows.
// Prints "GROUP OK" and exits 0 on success.
#include <unistd.h>
#include <sys/wait.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "apgk.h"

#define REQUIRE(c) do { if (!(c)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)
#define OK(call) do { int rc__ = (call); if (rc__ != APGK_OK) { std::fprintf(stderr, "FAILED %s:%d: %s -> %d\n", __FILE__, __LINE__, #call, rc__); return 1; } } while (0)

static const int K = 25;
static apgk_synth_params synth(uint64_t genome) {
  apgk_synth_params p{};
  p.genome_len = genome; p.seed_g = 0xA11BA7C5; p.seed_p = 0x5EED00A0; p.seed_q = 0x5EED00A1; p.seed_r = 0x5EED0001; p.seed_e = 0x5EED0002;
  p.read_len = 100; p.err_per_200 = 1;
  return p;
}
static int make_ctx(int device, apgk_ctx** c) {
  apgk_config cfg{};
  cfg.K = K; cfg.device = device; cfg.flags = APGK_WANT_COUNTS;
  return apgk_create(&cfg, c);
}
struct Table { std::vector<uint64_t> k; std::vector<uint32_t> c; uint64_t ni = 0, nd = 0; std::vector<uint64_t> f, m; };

// the single-GPU answer over reads [0, n_total)
static int reference_answer(int device, uint64_t genome, uint64_t n_total, Table& t) {
  apgk_ctx* c = nullptr;
  OK(make_ctx(device, &c));
  apgk_synth_params sp = synth(genome);
  OK(apgk_synth_reads(c, &sp, 0, n_total));
  OK(apgk_finish(c));
  OK(apgk_totals(c, &t.ni, &t.nd));
  t.k.resize(t.nd); t.c.resize(t.nd);
  OK(apgk_counts_copy(c, 0, t.nd, t.k.data(), t.c.data()));
  const uint64_t *f, *m; uint64_t n;
  OK(apgk_spectrum_sparse(c, &f, &m, &n));
  t.f.assign(f, f + n); t.m.assign(m, m + n);
  apgk_destroy(c);
  return 0;
}
// a shard's table must be a sub-sequence of the reference table with the same counts
static int check_shard(apgk_ctx* c, const Table& ref, uint64_t* n_found) {
  uint64_t ni, nd;
  OK(apgk_totals(c, &ni, &nd));
  std::vector<uint64_t> k(nd); std::vector<uint32_t> cnt(nd);
  if (nd) OK(apgk_counts_copy(c, 0, nd, k.data(), cnt.data()));
  uint64_t sum = 0;
  for (uint64_t i = 0; i < nd; i++) {
    REQUIRE(i == 0 || k[i - 1] < k[i]);
    const auto it = std::lower_bound(ref.k.begin(), ref.k.end(), k[i]);
    REQUIRE(it != ref.k.end() && *it == k[i]);
    REQUIRE(ref.c[it - ref.k.begin()] == cnt[i]);
    sum += cnt[i];
  }
  REQUIRE(sum == ni);
  *n_found = nd;
  return 0;
}
static int check_global(apgk_group* g, const Table& ref) {
  uint64_t ni, nd;
  OK(apgk_group_totals(g, &ni, &nd));
  REQUIRE(ni == ref.ni && nd == ref.nd);
  const uint64_t *f, *m; uint64_t n;
  OK(apgk_group_spectrum_sparse(g, &f, &m, &n));
  REQUIRE(n == ref.f.size());
  for (uint64_t i = 0; i < n; i++) REQUIRE(f[i] == ref.f[i] && m[i] == ref.m[i]);
  return 0;
}

static int run_local(int world) {
  const uint64_t genome = 400000;
  for (uint64_t n_per : {20000ull, 50000ull}) {
    std::vector<apgk_ctx*> cs(world, nullptr);
    apgk_synth_params sp = synth(genome);
    for (int r = 0; r < world; r++) { OK(make_ctx(0, &cs[r])); OK(apgk_synth_reads(cs[r], &sp, r * n_per, n_per)); }
    apgk_group* g = nullptr;
    OK(apgk_group_local(cs.data(), world, &g));
    Table ref;
    if (reference_answer(0, genome, n_per * world, ref)) return 1;
    for (int step = 0; step < 2; step++) {
      if (apgk_group_count(g) != APGK_OK) { std::fprintf(stderr, "apgk_group_count: %s\n", apgk_group_last_error(g)); return 1; }
      if (check_global(g, ref)) return 1;
      uint64_t total = 0;
      for (int r = 0; r < world; r++) { uint64_t nf = 0; if (check_shard(cs[r], ref, &nf)) return 1; total += nf; }
      REQUIRE(total == ref.nd);
    }
    apgk_group_stats st;
    OK(apgk_group_stats_get(g, &st));
    REQUIRE(st.world == world && st.n_rounds >= 1);
    // either order of destruction is fine: contexts first here
    apgk_destroy(cs[0]);
    apgk_group_destroy(g);
    for (int r = 1; r < world; r++) apgk_destroy(cs[r]);
  }
  std::printf("GROUP OK (single process, %d contexts on one device)\n", world);
  return 0;
}

static int rank_main(int rank, int world, const uint8_t* id) {
  const uint64_t genome = 2000000, n_per = 300000;
  apgk_ctx* c = nullptr;
  OK(make_ctx(rank, &c));
  apgk_synth_params sp = synth(genome);
  OK(apgk_synth_reads(c, &sp, rank * n_per, n_per));
  apgk_group* g = nullptr;
  if (apgk_group_join(c, id, rank, world, &g) != APGK_OK) { std::fprintf(stderr, "rank %d: join: %s\n", rank, apgk_last_error(c)); return 1; }
  Table ref;
  if (reference_answer(rank, genome, n_per * world, ref)) return 1;   // every rank computes the whole answer on its own GPU
  for (int step = 0; step < 2; step++) {
    if (apgk_group_count(g) != APGK_OK) { std::fprintf(stderr, "rank %d: count: %s\n", rank, apgk_group_last_error(g)); return 1; }
    if (check_global(g, ref)) return 1;
    uint64_t nf = 0;
    if (check_shard(c, ref, &nf)) return 1;
  }
  apgk_group_stats st;
  OK(apgk_group_stats_get(g, &st));
  if (rank == 0) std::printf("rank 0: shard %llu instances, %.1f MB over NVLink in %.3f ms, step %.2f ms\n", (unsigned long long)st.shard_instances,
                             st.remote_bytes / 1e6, st.gather_ms, st.step_ms);
  apgk_group_destroy(g);
  apgk_destroy(c);
  return 0;
}

static int run_procs(int world) {
  // count the GPUs in a child: the parent must not hold a CUDA context across fork()
  int pfd[2];
  if (pipe(pfd)) return 1;
  pid_t probe = fork();
  if (probe == 0) {
    apgk_ctx* c = nullptr;
    int n = 0;
    while (n < 64 && make_ctx(n, &c) == APGK_OK) { apgk_destroy(c); n++; }
    if (write(pfd[1], &n, sizeof n) != (ssize_t)sizeof n) _exit(1);
    _exit(0);
  }
  int ngpu = 0;
  if (read(pfd[0], &ngpu, sizeof ngpu) != (ssize_t)sizeof ngpu) ngpu = 0;
  waitpid(probe, nullptr, 0);
  if (ngpu < world) { std::printf("SKIPPED: %d GPUs, %d needed\n", ngpu, world); return 77; }
  // rank 0 makes the id and hands it to the others through pipes
  std::vector<int> rd(world, -1), wr(world, -1);
  for (int r = 1; r < world; r++) { int p[2]; if (pipe(p)) return 1; rd[r] = p[0]; wr[r] = p[1]; }
  std::vector<pid_t> kids;
  for (int r = 0; r < world; r++) {
    pid_t pid = fork();
    if (pid == 0) {
      uint8_t id[APGK_GROUP_ID_BYTES];
      if (r == 0) {
        if (apgk_group_unique_id(id) != APGK_OK) { std::fprintf(stderr, "apgk_group_unique_id failed (libnccl.so.2?)\n"); _exit(1); }
        for (int s = 1; s < world; s++) if (write(wr[s], id, sizeof id) != (ssize_t)sizeof id) _exit(1);
      } else if (read(rd[r], id, sizeof id) != (ssize_t)sizeof id) _exit(1);
      _exit(rank_main(r, world, id));
    }
    kids.push_back(pid);
  }
  int bad = 0;
  for (pid_t k : kids) { int st = 0; waitpid(k, &st, 0); if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) bad++; }
  if (bad) { std::fprintf(stderr, "%d rank(s) failed\n", bad); return 1; }
  std::printf("GROUP OK (%d processes, one GPU each, NCCL + CUDA IPC)\n", world);
  return 0;
}

int main(int argc, char** argv) {
  const bool procs = argc > 1 && !std::strcmp(argv[1], "procs");
  const int world = argc > 2 ? std::atoi(argv[2]) : 3;
  if (world < 1 || world > 16) return 2;
  return procs ? run_procs(world) : run_local(world);
}
