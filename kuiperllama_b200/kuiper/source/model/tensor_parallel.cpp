// Tensor-parallel plumbing of the C++ host side: configuration from the environment, the shard
// rules, and a TCP rendezvous for the few hundred bytes the ranks have to exchange at start-up
// (see model/tensor_parallel.h).  Plain POSIX sockets; nothing here touches the GPU.
#include "model/tensor_parallel.h"

#include <arpa/inet.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <sys/socket.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <thread>

namespace model {
namespace {
int env_int(const char* name, int fallback) {
  const char* v = std::getenv(name);
  return (v && *v) ? std::atoi(v) : fallback;
}

bool send_all(int fd, const void* p, size_t n) {
  const char* c = static_cast<const char*>(p);
  while (n > 0) {
    const ssize_t k = ::send(fd, c, n, MSG_NOSIGNAL);
    if (k < 0 && errno == EINTR) continue;
    if (k <= 0) return false;
    c += k, n -= static_cast<size_t>(k);
  }
  return true;
}
bool recv_all(int fd, void* p, size_t n) {
  char* c = static_cast<char*>(p);
  while (n > 0) {
    const ssize_t k = ::recv(fd, c, n, 0);
    if (k < 0 && errno == EINTR) continue;
    if (k <= 0) return false;
    c += k, n -= static_cast<size_t>(k);
  }
  return true;
}
}  // namespace

TpConfig TpConfig::from_env() {
  TpConfig c;
  c.world = std::max(1, env_int("KUIPER_TP_WORLD", 1));
  c.rank = env_int("KUIPER_TP_RANK", 0);
  c.device = env_int("KUIPER_TP_DEVICE", -1);
  if (const char* a = std::getenv("KUIPER_TP_ADDR"); a && *a) c.addr = a;
  c.port = env_int("KUIPER_TP_PORT", c.port);
  return c;
}

// Same rules as kuiperllama_b200/tensor_parallel.py (kv_heads_of_rank, ffn_range, check_shardable).
base::Status tp_shard(const TransformerConfig& c, int32_t group_size, int world, int rank, TpShard* out) {
  using base::error::InvalidArgument;
  if (world < 1 || rank < 0 || rank >= world) return InvalidArgument("tensor parallel: rank outside the world");
  if (world != 1 && world != 2 && world != 4 && world != 8)
    return InvalidArgument("tensor parallel: the exchange is built for 1, 2, 4 or 8 ranks");
  const int32_t hs = c.head_size_, heads = c.head_num_, kvh = c.kv_head_num_, hid = c.hidden_dim_;
  if (heads % world) return InvalidArgument("tensor parallel: the query heads do not split over the ranks");
  TpShard s;
  s.head_num = heads / world;
  s.q0 = rank * s.head_num * hs, s.q1 = (rank + 1) * s.head_num * hs;
  if (kvh >= world) {
    if (kvh % world) return InvalidArgument("tensor parallel: the kv heads do not split over the ranks");
    s.kv_head_num = kvh / world;
    s.k0 = rank * s.kv_head_num * hs;
  } else {
    // fewer kv heads than ranks (TinyLlama at 8): the ranks that share a kv head each keep a copy
    if (world % kvh || s.head_num > c.kv_mul_ || c.kv_mul_ % s.head_num)
      return InvalidArgument("tensor parallel: cannot replicate the kv heads over the ranks");
    s.kv_head_num = 1;
    s.k0 = (rank * s.head_num / c.kv_mul_) * hs;
  }
  s.k1 = s.k0 + s.kv_head_num * hs;
  if (group_size > 0) {
    if (hid % kTpInt8FfnUnit || hid / kTpInt8FfnUnit < world)
      return InvalidArgument("tensor parallel: the int8 FFN does not split in units of 256 columns");
    if ((s.head_num * hs) % group_size)
      return InvalidArgument("tensor parallel: quantisation groups straddle the split of the attention columns");
    const int32_t units = hid / kTpInt8FfnUnit, base_units = units / world, extra = units % world;
    const int32_t start = rank * base_units + std::min(rank, extra);
    s.f0 = start * kTpInt8FfnUnit;
    s.f1 = (start + base_units + (rank < extra ? 1 : 0)) * kTpInt8FfnUnit;
  } else {
    if (hid % world || (hid / world) % 4)
      return InvalidArgument("tensor parallel: hidden_dim / world must be a multiple of 4");
    s.f0 = rank * (hid / world), s.f1 = (rank + 1) * (hid / world);
  }
  s.hidden_dim = s.f1 - s.f0;
  *out = s;
  return base::error::Success();
}

int32_t tp_comm_words(const TransformerConfig& c, int world) {
  const int32_t per_rank = (c.vocab_size_ % std::max(world, 1) == 0) ? c.vocab_size_ / std::max(world, 1) : 0;
  return (std::max(c.dim_, per_rank) + 3) / 4 * 4;
}

// ---- rendezvous ----------------------------------------------------------------------------------------
TpRendezvous::~TpRendezvous() { close(); }

void TpRendezvous::close() {
  for (int fd : peers_)
    if (fd >= 0) ::close(fd);
  peers_.clear();
  if (listen_fd_ >= 0) ::close(listen_fd_);
  listen_fd_ = -1;
  open_ = false;
}

base::Status TpRendezvous::open(const TpConfig& cfg, int timeout_s) {
  using base::error::InternalError;
  close();
  world_ = cfg.world, rank_ = cfg.rank;
  if (world_ <= 1) {
    open_ = true;
    return base::error::Success();
  }
  sockaddr_in sa{};
  sa.sin_family = AF_INET;
  sa.sin_port = htons(static_cast<uint16_t>(cfg.port));
  if (::inet_pton(AF_INET, cfg.addr.c_str(), &sa.sin_addr) != 1)
    return InternalError("tensor parallel rendezvous: '" + cfg.addr + "' is not an IPv4 address");
  const int one = 1;
  const auto deadline = std::chrono::steady_clock::now() + std::chrono::seconds(timeout_s);
  if (rank_ == 0) {
    listen_fd_ = ::socket(AF_INET, SOCK_STREAM, 0);
    if (listen_fd_ < 0) return InternalError("tensor parallel rendezvous: socket() failed");
    ::setsockopt(listen_fd_, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
    if (::bind(listen_fd_, reinterpret_cast<sockaddr*>(&sa), sizeof(sa)) != 0 || ::listen(listen_fd_, world_) != 0)
      return InternalError("tensor parallel rendezvous: cannot listen on " + cfg.addr + ":" +
                           std::to_string(cfg.port) + " (" + std::strerror(errno) + ")");
    timeval tv{timeout_s, 0};
    ::setsockopt(listen_fd_, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));  // accept() gives up eventually
    peers_.assign(world_, -1);
    for (int got = 0; got < world_ - 1; ++got) {
      const int fd = ::accept(listen_fd_, nullptr, nullptr);
      if (fd < 0) return InternalError("tensor parallel rendezvous: a rank did not show up");
      ::setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
      const timeval patience{10 * timeout_s, 0};  // a peer that stops answering ends in an error, not a hang
      ::setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &patience, sizeof(patience));
      int32_t who = -1;
      if (!recv_all(fd, &who, sizeof(who)) || who <= 0 || who >= world_ || peers_[who] >= 0) {
        ::close(fd);
        return InternalError("tensor parallel rendezvous: bad hello from a peer");
      }
      peers_[who] = fd;
    }
  } else {
    int fd = -1;
    for (;;) {
      fd = ::socket(AF_INET, SOCK_STREAM, 0);
      if (fd < 0) return InternalError("tensor parallel rendezvous: socket() failed");
      if (::connect(fd, reinterpret_cast<sockaddr*>(&sa), sizeof(sa)) == 0) break;
      ::close(fd);
      if (std::chrono::steady_clock::now() > deadline)
        return InternalError("tensor parallel rendezvous: rank 0 is not listening on " + cfg.addr + ":" +
                             std::to_string(cfg.port));
      std::this_thread::sleep_for(std::chrono::milliseconds(50));
    }
    ::setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
    const timeval patience{10 * timeout_s, 0};
    ::setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &patience, sizeof(patience));
    const int32_t who = rank_;
    if (!send_all(fd, &who, sizeof(who))) {
      ::close(fd);
      return InternalError("tensor parallel rendezvous: hello failed");
    }
    peers_.assign(1, fd);
  }
  open_ = true;
  return base::error::Success();
}

base::Status TpRendezvous::all_gather(const void* mine, size_t bytes, void* all) {
  using base::error::InternalError;
  if (!open_) return InternalError("tensor parallel rendezvous: not open");
  char* dst = static_cast<char*>(all);
  if (world_ <= 1) {
    std::memcpy(dst, mine, bytes);
    return base::error::Success();
  }
  if (rank_ == 0) {
    std::memcpy(dst, mine, bytes);
    for (int r = 1; r < world_; ++r)
      if (!recv_all(peers_[r], dst + static_cast<size_t>(r) * bytes, bytes))
        return InternalError("tensor parallel rendezvous: lost rank " + std::to_string(r));
    for (int r = 1; r < world_; ++r)
      if (!send_all(peers_[r], dst, bytes * static_cast<size_t>(world_)))
        return InternalError("tensor parallel rendezvous: lost rank " + std::to_string(r));
  } else {
    if (!send_all(peers_[0], mine, bytes) || !recv_all(peers_[0], dst, bytes * static_cast<size_t>(world_)))
      return InternalError("tensor parallel rendezvous: lost rank 0");
  }
  return base::error::Success();
}

base::Status TpRendezvous::barrier() {
  std::vector<char> all(static_cast<size_t>(std::max(world_, 1)));
  const char me = 1;
  return all_gather(&me, 1, all.data());
}
}  // namespace model
