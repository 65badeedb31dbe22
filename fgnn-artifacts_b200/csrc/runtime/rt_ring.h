// Ticket protocol of the arch5 sampler -> trainer queue (reference: MemoryQueue, memory_queue.cc:51-138 and
// memory_queue.h:46-113: a bounded ring of fixed-size slots in shared memory, semaphores for the fill level).
// Several samplers produce and several trainers consume; writers finish out of order (a slow sampler is still
// copying ticket t while ticket t+1 is already published) and readers release out of order (a slow trainer keeps
// its slot while a fast one returns the next).  The fill-level semaphores therefore only provide back-pressure;
// what makes a slot safe to touch is its sequence word (the bounded MPMC queue of D. Vyukov):
//     seq == t          slot is free for the writer holding ticket t        (initially seq = slot index)
//     seq == t + 1      ticket t's record is complete, its reader may copy it out
//     seq == t + N      the reader is done, the writer of ticket t + N may start   (N = num_slots)
// A plain free/published flag is NOT enough: "free" would also be what the writer of ticket t + N sees while the
// writer of ticket t has claimed the slot but not published yet.  Works across processes (MAP_SHARED,
// process-shared mutex/semaphores) and across threads (tests/test_ring_cpu.py drives fgnn_rt_ring_selftest).
#pragma once
#include <pthread.h>
#include <semaphore.h>
#include <stdint.h>

#include <atomic>
#include <chrono>
#include <thread>

namespace fgnn {
namespace rt {

struct RingCtl {
  pthread_mutex_t mu;
  sem_t free_slots, used_slots;
  uint64_t head, tail;
  uint32_t num_slots;
};

inline void RingInit(RingCtl *c, uint32_t num_slots, bool process_shared) {
  pthread_mutexattr_t ma;
  pthread_mutexattr_init(&ma);
  if (process_shared) pthread_mutexattr_setpshared(&ma, PTHREAD_PROCESS_SHARED);
  pthread_mutex_init(&c->mu, &ma);
  pthread_mutexattr_destroy(&ma);
  sem_init(&c->free_slots, process_shared ? 1 : 0, num_slots);
  sem_init(&c->used_slots, process_shared ? 1 : 0, 0);
  c->head = c->tail = 0;
  c->num_slots = num_slots;
}

inline void RingPause() { std::this_thread::sleep_for(std::chrono::microseconds(1)); }

inline void RingInitSlot(std::atomic<uint32_t> *seq, uint32_t slot_index) { seq->store(slot_index); }

// Producer: take the next ticket and wait until its slot has been released by the reader of ticket - num_slots.
// `seq_of(ticket)` returns the slot's sequence word.  Returns false when `stop` was raised while waiting (nothing
// may be written then).
template <typename SeqOf>
inline bool RingBeginWrite(RingCtl *c, SeqOf seq_of, const std::atomic<bool> *stop, uint64_t *ticket,
                           bool wait_for_slot = true /* false only in the self-test's negative control */) {
  while (sem_trywait(&c->free_slots) != 0) {
    if (stop && stop->load(std::memory_order_relaxed)) return false;
    RingPause();
  }
  pthread_mutex_lock(&c->mu);
  *ticket = c->tail++;
  pthread_mutex_unlock(&c->mu);
  std::atomic<uint32_t> *seq = seq_of(*ticket);
  while (wait_for_slot && seq->load(std::memory_order_acquire) != (uint32_t)*ticket) {
    if (stop && stop->load(std::memory_order_relaxed)) return false;
    RingPause();
  }
  return true;
}
inline void RingEndWrite(RingCtl *c, std::atomic<uint32_t> *seq, uint64_t ticket) {
  seq->store((uint32_t)(ticket + 1), std::memory_order_release);
  sem_post(&c->used_slots);
}

// Consumer: take the next ticket (optionally without blocking when the ring is empty) and wait until its record
// has been published — with several producers a later ticket can be complete before an earlier one.
template <typename SeqOf>
inline bool RingBeginRead(RingCtl *c, SeqOf seq_of, const std::atomic<bool> *stop, bool block, uint64_t *ticket) {
  while (sem_trywait(&c->used_slots) != 0) {
    if (!block || (stop && stop->load(std::memory_order_relaxed))) return false;
    RingPause();
  }
  pthread_mutex_lock(&c->mu);
  *ticket = c->head++;
  pthread_mutex_unlock(&c->mu);
  std::atomic<uint32_t> *seq = seq_of(*ticket);
  while (seq->load(std::memory_order_acquire) != (uint32_t)(*ticket + 1)) {
    if (stop && stop->load(std::memory_order_relaxed)) return false;
    RingPause();
  }
  return true;
}
inline void RingEndRead(RingCtl *c, std::atomic<uint32_t> *seq, uint64_t ticket) {
  seq->store((uint32_t)(ticket + c->num_slots), std::memory_order_release);
  sem_post(&c->free_slots);
}

}  // namespace rt
}  // namespace fgnn

// SAMGRAPH_SANITY_CHECK=1 (reference: cuda_shuffler.cc:144-151 -> GPUSanityCheckList / GPUBatchSanityCheck,
// cuda_sanity_check.cu:29-59): every mini-batch's seeds must be valid ids and no training node may be handed out
// twice within an epoch.  `epoch_map` has one byte per vertex, cleared by the caller at every epoch start.
// Returns 0 = ok, 1 = an id equals the empty key, 2 = id out of range, 3 = "duplicate batch input"; *bad_index
// (optional) receives the position of the offending seed.  Host function: the engine copies the <= batch_size
// seeds back only when the check is switched on.
extern "C" int fgnn_rt_sanity_check_batch(uint8_t *epoch_map, size_t num_nodes, const uint32_t *seeds, size_t n,
                                          size_t *bad_index);

// Threaded stress test of the protocol above (test hook, CPU only): `producers` threads publish `items` records
// in total into a ring of `num_slots` slots of `slot_words` 32-bit words, `consumers` threads take them out,
// sleeping up to `max_delay_us` (pseudo-random per record) while they hold a slot so that slots are released out
// of order.  Every record carries its sequence number in every word.  Returns 0 when every record arrived exactly
// once and intact, a positive count of damaged / duplicated / missing records otherwise, -1 on timeout
// (`timeout_ms`): the consumers or producers stopped making progress.  `unsafe_no_slot_wait` != 0 is the negative
// control: producers rely on the fill-level semaphore alone — records get overwritten while a slow consumer still
// holds them, and the sequence words go out of step (dead-lock).  2 and 3 model a consumer whose slot release is
// deferred (asynchronous copies out of the slot, Engine::RecvTask): 3 releases the previous slot BEFORE asking for
// the next ticket (the engine's order), 2 only after it got one — with more consumers than slots the ticket it
// then waits for can map to the slot it still holds (second negative control: times out).
extern "C" long fgnn_rt_ring_selftest(uint32_t num_slots, uint32_t slot_words, uint32_t producers,
                                      uint32_t consumers, uint64_t items, uint32_t max_delay_us,
                                      uint32_t timeout_ms, int unsafe_no_slot_wait);
// The same stress test with PROCESSES: ring, slots and counters live in one MAP_SHARED region, producers and
// consumers are fork()ed children (the arch5 situation: process-shared mutex / semaphores, atomics in shared
// memory).  Call it from a single-threaded process.  Same return values; -3 = fork/mmap failure.
extern "C" long fgnn_rt_ring_selftest_procs(uint32_t num_slots, uint32_t slot_words, uint32_t producers,
                                            uint32_t consumers, uint64_t items, uint32_t max_delay_us,
                                            uint32_t timeout_ms);
