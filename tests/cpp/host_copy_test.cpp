// host_copy_test.cpp - HostCopier (serenity_b200/csrc/host_copy.h) without a GPU: the streamed form that staged_d2h drives
// (open with a first delivered piece, publish the mark piece by piece from a "DMA" thread, help() from the owner, close) has to
// copy every byte exactly once for every worker count, including zero workers and sizes that are no multiple of the slice.
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "../../serenity_b200/csrc/host_copy.h"

static int check(int nworkers, size_t bytes, size_t piece, bool delayed) {
  std::vector<unsigned char> staged(bytes, 0), dst(bytes + 64, 0xEE);
  sxc::HostCopier copier(nworkers);
  std::atomic<size_t> delivered{0};
  // the "DMA engine": fills the staging buffer piece by piece
  std::thread dma([&] {
    for (size_t off = 0; off < bytes; off += piece) {
      const size_t end = std::min(bytes, off + piece);
      for (size_t i = off; i < end; ++i) staged[i] = (unsigned char)((i * 2654435761u) >> 13);
      if (delayed) std::this_thread::sleep_for(std::chrono::microseconds(200));
      delivered.store(end, std::memory_order_release);
    }
  });
  while (delivered.load(std::memory_order_acquire) < std::min(piece, bytes)) std::this_thread::yield();
  copier.open(dst.data(), staged.data(), bytes, std::min(piece, bytes));
  size_t mark = std::min(piece, bytes);
  while (mark < bytes) {
    const size_t d = delivered.load(std::memory_order_acquire);
    if (d > mark) {
      mark = d;
      copier.publish(mark);
    } else if (!copier.help()) {
      std::this_thread::yield();
    }
  }
  copier.publish(bytes);
  copier.close();
  dma.join();
  for (size_t i = 0; i < bytes; ++i)
    if (dst[i] != (unsigned char)((i * 2654435761u) >> 13)) {
      std::printf("FAIL workers=%d bytes=%zu piece=%zu: byte %zu differs\n", nworkers, bytes, piece, i);
      return 1;
    }
  for (size_t i = bytes; i < bytes + 64; ++i)
    if (dst[i] != 0xEE) {
      std::printf("FAIL workers=%d bytes=%zu: wrote past the end\n", nworkers, bytes);
      return 1;
    }
  // the plain form on the same object afterwards (a job must leave the copier reusable)
  std::vector<unsigned char> again(bytes, 0);
  copier.copy(again.data(), staged.data(), bytes);
  if (again != staged) {
    std::printf("FAIL workers=%d bytes=%zu: plain copy after a streamed job differs\n", nworkers, bytes);
    return 1;
  }
  return 0;
}

int main() {
  int bad = 0;
  const size_t sizes[] = {1, 4096, (256u << 10) - 1, 256u << 10, (1u << 20) + 17, 3175200, (8u << 20) + 5};
  for (int w : {0, 1, 3, 7})
    for (size_t b : sizes)
      for (size_t piece : {(size_t)512 << 10, (size_t)100000}) bad += check(w, b, piece, b == 3175200 && w == 3);
  std::printf(bad ? "host_copy_test: %d FAILED\n" : "host_copy_test: ok\n", bad);
  return bad ? 1 : 0;
}
