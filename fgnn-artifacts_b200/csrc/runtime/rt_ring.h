// Ticket protocol of the arch5 sampler -> trainer queue (reference: MemoryQueue, memory_queue.cc:51-138 and
// memory_queue.h:46-113: a bounded ring of fixed-size slots in shared memory, semaphores for the fill level).
// Several samplers produce and several trainers consume, and trainers release their slots OUT OF ORDER (a slow
// trainer keeps its slot while a fast one returns the next), so the fill-level semaphore alone does not say that
// the slot a new ticket maps to (ticket % num_slots) is free: every slot carries a `ready` word
//     0 = free        (set by the consumer after it has copied the record out)
//     1 = published   (set by the producer after the record is complete)
// and both sides wait on the word of THEIR slot after taking a ticket.  Works across processes (MAP_SHARED,
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

// Producer: take the next ticket and wait until its slot has been released.  `ready_of(ticket)` returns the
// slot's ready word.  Returns false when `stop` was raised while waiting (nothing may be written then).
template <typename ReadyOf>
inline bool RingBeginWrite(RingCtl *c, ReadyOf ready_of, const std::atomic<bool> *stop, uint64_t *ticket,
                           bool wait_for_slot = true /* false only in the self-test's negative control */) {
  while (sem_trywait(&c->free_slots) != 0) {
    if (stop && stop->load(std::memory_order_relaxed)) return false;
    RingPause();
  }
  pthread_mutex_lock(&c->mu);
  *ticket = c->tail++;
  pthread_mutex_unlock(&c->mu);
  std::atomic<uint32_t> *ready = ready_of(*ticket);
  while (wait_for_slot && ready->load(std::memory_order_acquire) != 0) {  // its previous reader is still copying it out
    if (stop && stop->load(std::memory_order_relaxed)) return false;
    RingPause();
  }
  return true;
}
inline void RingEndWrite(RingCtl *c, std::atomic<uint32_t> *ready) {
  ready->store(1, std::memory_order_release);
  sem_post(&c->used_slots);
}

// Consumer: take the next ticket (optionally without blocking when the ring is empty) and wait until its record
// has been published — with several producers a later ticket can be complete before an earlier one.
template <typename ReadyOf>
inline bool RingBeginRead(RingCtl *c, ReadyOf ready_of, const std::atomic<bool> *stop, bool block, uint64_t *ticket) {
  while (sem_trywait(&c->used_slots) != 0) {
    if (!block || (stop && stop->load(std::memory_order_relaxed))) return false;
    RingPause();
  }
  pthread_mutex_lock(&c->mu);
  *ticket = c->head++;
  pthread_mutex_unlock(&c->mu);
  std::atomic<uint32_t> *ready = ready_of(*ticket);
  while (ready->load(std::memory_order_acquire) == 0) {
    if (stop && stop->load(std::memory_order_relaxed)) return false;
    RingPause();
  }
  return true;
}
inline void RingEndRead(RingCtl *c, std::atomic<uint32_t> *ready) {
  ready->store(0, std::memory_order_release);
  sem_post(&c->free_slots);
}

}  // namespace rt
}  // namespace fgnn

// Threaded stress test of the protocol above (test hook, CPU only): `producers` threads publish `items` records
// in total into a ring of `num_slots` slots of `slot_words` 32-bit words, `consumers` threads take them out,
// sleeping up to `max_delay_us` (pseudo-random per record) while they hold a slot so that slots are released out
// of order.  Every record carries its sequence number in every word.  Returns 0 when every record arrived exactly
// once and intact, a positive count of damaged / duplicated / missing records otherwise, -1 on timeout
// (`timeout_ms`): the consumers or producers stopped making progress.  `unsafe_no_slot_wait` != 0 is the negative
// control: producers rely on the fill-level semaphore alone (the round-1 bug) — records get overwritten while a
// slow consumer still holds them, or a slot's ready word is cleared after its next record was published.
extern "C" long fgnn_rt_ring_selftest(uint32_t num_slots, uint32_t slot_words, uint32_t producers,
                                      uint32_t consumers, uint64_t items, uint32_t max_delay_us,
                                      uint32_t timeout_ms, int unsafe_no_slot_wait);
