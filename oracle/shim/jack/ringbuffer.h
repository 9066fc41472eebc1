// TEST INFRASTRUCTURE: the JACK lock-free ring buffer API (jack/ringbuffer.h), restated from its
// documented semantics: capacity is rounded up to a power of two, one byte is kept free, and
// get_read_vector() exposes the readable region as at most two contiguous chunks.
#pragma once
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
typedef struct {
  char* buf;
  size_t len;
} jack_ringbuffer_data_t;
typedef struct {
  char* buf;
  volatile size_t write_ptr;
  volatile size_t read_ptr;
  size_t size;
  size_t size_mask;
  int mlocked;
} jack_ringbuffer_t;
static inline jack_ringbuffer_t* jack_ringbuffer_create(size_t sz) {
  jack_ringbuffer_t* rb = (jack_ringbuffer_t*)malloc(sizeof(jack_ringbuffer_t));
  int p2;
  for (p2 = 1; ((size_t)1 << p2) < sz; p2++) {}
  rb->size = (size_t)1 << p2;
  rb->size_mask = rb->size - 1;
  rb->write_ptr = rb->read_ptr = 0;
  rb->buf = (char*)malloc(rb->size);
  rb->mlocked = 0;
  return rb;
}
static inline void jack_ringbuffer_free(jack_ringbuffer_t* rb) { free(rb->buf); free(rb); }
static inline void jack_ringbuffer_reset(jack_ringbuffer_t* rb) { rb->read_ptr = rb->write_ptr = 0; }
static inline size_t jack_ringbuffer_read_space(const jack_ringbuffer_t* rb) { return (rb->write_ptr - rb->read_ptr) & rb->size_mask; }
static inline size_t jack_ringbuffer_write_space(const jack_ringbuffer_t* rb) {
  size_t w = rb->write_ptr, r = rb->read_ptr;
  if (w > r) return ((r - w + rb->size) & rb->size_mask) - 1;
  if (w < r) return (r - w) - 1;
  return rb->size - 1;
}
static inline size_t jack_ringbuffer_write(jack_ringbuffer_t* rb, const char* src, size_t cnt) {
  size_t free_cnt = jack_ringbuffer_write_space(rb);
  if (free_cnt == 0) return 0;
  size_t to_write = cnt > free_cnt ? free_cnt : cnt;
  size_t cnt2 = rb->write_ptr + to_write, n1, n2;
  if (cnt2 > rb->size) { n1 = rb->size - rb->write_ptr; n2 = cnt2 & rb->size_mask; } else { n1 = to_write; n2 = 0; }
  memcpy(rb->buf + rb->write_ptr, src, n1);
  rb->write_ptr = (rb->write_ptr + n1) & rb->size_mask;
  if (n2) { memcpy(rb->buf + rb->write_ptr, src + n1, n2); rb->write_ptr = (rb->write_ptr + n2) & rb->size_mask; }
  return to_write;
}
static inline void jack_ringbuffer_read_advance(jack_ringbuffer_t* rb, size_t cnt) { rb->read_ptr = (rb->read_ptr + cnt) & rb->size_mask; }
static inline void jack_ringbuffer_get_read_vector(const jack_ringbuffer_t* rb, jack_ringbuffer_data_t* vec) {
  size_t w = rb->write_ptr, r = rb->read_ptr;
  size_t free_cnt = (w - r) & rb->size_mask;
  size_t cnt2 = r + free_cnt;
  if (cnt2 > rb->size) {
    vec[0].buf = rb->buf + r; vec[0].len = rb->size - r;
    vec[1].buf = rb->buf;     vec[1].len = cnt2 & rb->size_mask;
  } else {
    vec[0].buf = rb->buf + r; vec[0].len = free_cnt;
    vec[1].buf = rb->buf;     vec[1].len = 0;
  }
}
