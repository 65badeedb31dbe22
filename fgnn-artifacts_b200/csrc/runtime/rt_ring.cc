// fgnn_rt_ring_selftest: threaded stress test of the arch5 queue's ticket protocol (rt_ring.h).
#include "rt_ring.h"

#include <memory>
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
        if (!RingBeginWrite(&ctl, ready_of, &stop, &ticket, unsafe_no_slot_wait == 0)) return;
        Slot &s = slots[ticket % num_slots];
        if ((ticket % 3) == p % 3) delay(ticket * 17 + p + 1000);  // slow writers: records complete out of order
        for (uint32_t w = 0; w < slot_words; ++w) {  // a slow, word-by-word write like a DMA in flight
          s.words[w] = (uint32_t)seq;
          if ((w & 63u) == 63u) std::this_thread::yield();
        }
        RingEndWrite(&ctl, &s.ready, ticket);
      }
    });
  for (uint32_t c = 0; c < consumers; ++c)
    threads.emplace_back([&, c] {
      for (;;) {
        if (consumed.load() >= items) return;
        uint64_t ticket;
        if (!RingBeginRead(&ctl, ready_of, &stop, /*block=*/false, &ticket)) {
          if (stop) return;
          RingPause();
          continue;
        }
        Slot &s = slots[ticket % num_slots];
        const uint32_t seq = s.words[0];
        delay(ticket * 131 + c);  // hold the slot for a while: releases happen out of order
        bool ok = seq < items;
        for (uint32_t w = 0; w < slot_words; ++w) ok &= (s.words[w] == seq);
        if (!ok) damaged.fetch_add(1);
        else seen[seq].fetch_add(1);
        RingEndRead(&ctl, &s.ready, ticket);
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
