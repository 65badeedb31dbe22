// fgnn_rt_ring_selftest: threaded stress test of the arch5 queue's ticket protocol (rt_ring.h).
#include "rt_ring.h"

#include <signal.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#include <memory>
#include <new>
#include <vector>

using namespace fgnn::rt;

extern "C" long fgnn_rt_ring_selftest(uint32_t num_slots, uint32_t slot_words, uint32_t producers,
                                      uint32_t consumers, uint64_t items, uint32_t max_delay_us,
                                      uint32_t timeout_ms, int unsafe_no_slot_wait) {
  if (!num_slots || !slot_words || !producers || !consumers) return -2;
  struct Slot {
    std::atomic<uint32_t> ready{0};
    std::vector<uint32_t> words;
  };
  RingCtl ctl;
  RingInit(&ctl, num_slots, /*process_shared=*/false);
  std::vector<Slot> slots(num_slots);
  for (uint32_t i = 0; i < num_slots; ++i) {
    slots[i].words.assign(slot_words, 0xFFFFFFFFu);
    RingInitSlot(&slots[i].ready, i);
  }
  auto ready_of = [&](uint64_t ticket) { return &slots[ticket % num_slots].ready; };

  std::atomic<bool> stop{false};
  std::atomic<uint64_t> next_item{0}, consumed{0}, damaged{0};
  std::unique_ptr<std::atomic<uint32_t>[]> seen(new std::atomic<uint32_t>[items ? items : 1]);
  for (uint64_t i = 0; i < items; ++i) seen[i] = 0;
  auto delay = [&](uint64_t x) {
    if (!max_delay_us) return;
    x = (x + 0x9E3779B97F4A7C15ull) * 0xBF58476D1CE4E5B9ull;
    x ^= x >> 29;
    const uint32_t us = (uint32_t)(x % (max_delay_us + 1));
    if (us) std::this_thread::sleep_for(std::chrono::microseconds(us));
  };

  std::vector<std::thread> threads;
  for (uint32_t p = 0; p < producers; ++p)
    threads.emplace_back([&, p] {
      for (;;) {
        const uint64_t seq = next_item.fetch_add(1);
        if (seq >= items) return;
        uint64_t ticket;
        if (!RingBeginWrite(&ctl, ready_of, &stop, &ticket, unsafe_no_slot_wait != 1)) return;
        Slot &s = slots[ticket % num_slots];
        if ((ticket % 3) == p % 3) delay(ticket * 17 + p + 1000);  // slow writers: records complete out of order
        for (uint32_t w = 0; w < slot_words; ++w) {  // a slow, word-by-word write like a DMA in flight
          s.words[w] = (uint32_t)seq;
          if ((w & 63u) == 63u) std::this_thread::yield();
        }
        RingEndWrite(&ctl, &s.ready, ticket);
      }
    });
  // mode 2 / 3: the engine's trainer since round 2 — the copies out of a slot are asynchronous, so the slot is
  // released LATER than the record is taken: in mode 3 (what Engine::RecvTask does) before the next ticket is
  // requested, in mode 2 (the bug it had) only after the next ticket has been obtained.
  const bool deferred = unsafe_no_slot_wait == 2 || unsafe_no_slot_wait == 3;
  const bool release_after_take = unsafe_no_slot_wait == 2;
  for (uint32_t c = 0; c < consumers; ++c)
    threads.emplace_back([&, c] {
      bool have_pending = false;
      uint64_t pending = 0;
      auto release_pending = [&] {
        if (!have_pending) return;
        RingEndRead(&ctl, &slots[pending % num_slots].ready, pending);
        have_pending = false;
      };
      for (;;) {
        if (consumed.load() >= items) { release_pending(); return; }
        uint64_t ticket;
        if (deferred && !release_after_take) release_pending();
        if (!RingBeginRead(&ctl, ready_of, &stop, /*block=*/false, &ticket)) {
          if (stop) return;
          if (release_after_take) { delay(c + 7); release_pending(); }  // an idle trainer does get round to its poll
          else RingPause();
          continue;
        }
        Slot &s = slots[ticket % num_slots];
        const uint32_t seq = s.words[0];
        delay(ticket * 131 + c);  // hold the slot for a while: releases happen out of order
        bool ok = seq < items;
        for (uint32_t w = 0; w < slot_words; ++w) ok &= (s.words[w] == seq);
        if (!ok) damaged.fetch_add(1);
        else seen[seq].fetch_add(1);
        if (deferred) {
          release_pending();  // mode 2 reaches this with the previous slot still held
          pending = ticket;
          have_pending = true;
        } else {
          RingEndRead(&ctl, &s.ready, ticket);
        }
        consumed.fetch_add(1);
      }
    });

  const auto t0 = std::chrono::steady_clock::now();
  bool timed_out = false;
  while (consumed.load() < items) {
    std::this_thread::sleep_for(std::chrono::milliseconds(1));
    if (std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(timeout_ms)) {
      timed_out = true;
      break;
    }
  }
  stop = true;
  for (auto &t : threads) t.join();
  sem_destroy(&ctl.free_slots);
  sem_destroy(&ctl.used_slots);
  pthread_mutex_destroy(&ctl.mu);
  if (timed_out) return -1;
  long bad = (long)damaged.load();
  for (uint64_t i = 0; i < items; ++i) bad += (seen[i].load() != 1);
  return bad;
}

extern "C" int fgnn_rt_sanity_check_batch(uint8_t *epoch_map, size_t num_nodes, const uint32_t *seeds, size_t n,
                                          size_t *bad_index) {
  for (size_t i = 0; i < n; ++i) {
    const uint32_t v = seeds[i];
    int rc = 0;
    if (v == 0xFFFFFFFFu) rc = 1;            // Constant::kEmptyKey (list_sanity_check)
    else if (v >= num_nodes) rc = 2;
    else if (epoch_map[v]) rc = 3;           // batch_sanity_check: map[input[index]] must still be 0
    if (rc) {
      if (bad_index) *bad_index = i;
      return rc;
    }
    epoch_map[v] = 1;
  }
  return 0;
}

namespace {
struct ProcShared {
  RingCtl ctl;
  std::atomic<uint64_t> next_item, consumed, damaged;
};
inline void proc_delay(uint64_t x, uint32_t max_delay_us) {
  if (!max_delay_us) return;
  x = (x + 0x9E3779B97F4A7C15ull) * 0xBF58476D1CE4E5B9ull;
  x ^= x >> 29;
  const uint32_t us = (uint32_t)(x % (max_delay_us + 1));
  if (us) usleep(us);
}
}  // namespace

extern "C" long fgnn_rt_ring_selftest_procs(uint32_t num_slots, uint32_t slot_words, uint32_t producers,
                                            uint32_t consumers, uint64_t items, uint32_t max_delay_us,
                                            uint32_t timeout_ms) {
  if (!num_slots || !slot_words || !producers || !consumers) return -2;
  // layout: ProcShared | seq[num_slots] | seen[items] | words[num_slots][slot_words]
  const size_t off_seq = (sizeof(ProcShared) + 63) & ~(size_t)63;
  const size_t off_seen = off_seq + (((size_t)num_slots * sizeof(std::atomic<uint32_t>) + 63) & ~(size_t)63);
  const size_t off_words = off_seen + (((size_t)(items ? items : 1) * sizeof(std::atomic<uint32_t>) + 63) & ~(size_t)63);
  const size_t bytes = off_words + (size_t)num_slots * slot_words * sizeof(uint32_t);
  char *base = (char *)mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
  if (base == MAP_FAILED) return -3;
  ProcShared *sh = new (base) ProcShared();
  auto *seq = reinterpret_cast<std::atomic<uint32_t> *>(base + off_seq);
  auto *seen = reinterpret_cast<std::atomic<uint32_t> *>(base + off_seen);
  auto *words = reinterpret_cast<volatile uint32_t *>(base + off_words);
  RingInit(&sh->ctl, num_slots, /*process_shared=*/true);
  for (uint32_t i = 0; i < num_slots; ++i) RingInitSlot(new (&seq[i]) std::atomic<uint32_t>(0), i);
  for (uint64_t i = 0; i < items; ++i) new (&seen[i]) std::atomic<uint32_t>(0);
  sh->next_item = 0;
  sh->consumed = 0;
  sh->damaged = 0;
  auto seq_of = [&](uint64_t ticket) { return &seq[ticket % num_slots]; };

  std::vector<pid_t> kids;
  bool fork_failed = false;
  for (uint32_t k = 0; k < producers + consumers && !fork_failed; ++k) {
    const pid_t pid = fork();
    if (pid < 0) { fork_failed = true; break; }
    if (pid == 0) {  // child: no allocation, no stdio — only the ring
      if (k < producers) {
        const uint32_t p = k;
        for (;;) {
          const uint64_t item = sh->next_item.fetch_add(1);
          if (item >= items) _exit(0);
          uint64_t ticket;
          if (!RingBeginWrite(&sh->ctl, seq_of, nullptr, &ticket)) _exit(0);
          volatile uint32_t *w = words + (size_t)(ticket % num_slots) * slot_words;
          if ((ticket % 3) == p % 3) proc_delay(ticket * 17 + p + 1000, max_delay_us);
          for (uint32_t i = 0; i < slot_words; ++i) w[i] = (uint32_t)item;
          RingEndWrite(&sh->ctl, seq_of(ticket), ticket);
        }
      } else {
        const uint32_t c = k - producers;
        for (;;) {
          if (sh->consumed.load() >= items) _exit(0);
          uint64_t ticket;
          if (!RingBeginRead(&sh->ctl, seq_of, nullptr, /*block=*/false, &ticket)) {
            usleep(1);
            continue;
          }
          volatile uint32_t *w = words + (size_t)(ticket % num_slots) * slot_words;
          const uint32_t item = w[0];
          proc_delay(ticket * 131 + c, max_delay_us);
          bool ok = item < items;
          for (uint32_t i = 0; i < slot_words; ++i) ok &= (w[i] == item);
          if (!ok) sh->damaged.fetch_add(1);
          else seen[item].fetch_add(1);
          RingEndRead(&sh->ctl, seq_of(ticket), ticket);
          sh->consumed.fetch_add(1);
        }
      }
    }
    kids.push_back(pid);
  }
  const auto t0 = std::chrono::steady_clock::now();
  bool timed_out = false;
  while (!fork_failed && sh->consumed.load() < items) {
    usleep(1000);
    if (std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(timeout_ms)) {
      timed_out = true;
      break;
    }
  }
  if (timed_out || fork_failed)
    for (pid_t pid : kids) kill(pid, SIGKILL);
  else
    usleep(20000);  // consumers notice consumed >= items; producers are blocked on a full-ring wait or gone
  for (pid_t pid : kids) {
    if (!(timed_out || fork_failed)) kill(pid, SIGKILL);  // producers waiting for a ticket nobody will free
    int st = 0;
    waitpid(pid, &st, 0);
  }
  long result;
  if (fork_failed) result = -3;
  else if (timed_out) result = -1;
  else {
    result = (long)sh->damaged.load();
    for (uint64_t i = 0; i < items; ++i) result += (seen[i].load() != 1);
  }
  munmap(base, bytes);
  return result;
}
