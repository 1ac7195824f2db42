// pose_mailbox.cpp -- see pose_mailbox.hpp
#include <pose_mailbox.hpp>
#include <chrono>
#include <cstring>
#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace kf
{
namespace
{
inline void relax()
{
#if defined(__x86_64__)
    _mm_pause();
#endif
}
} // namespace

bool PoseMailbox::open(const std::string &name, int rank, int world)
{
    close();
    if (world < 1 || world > 64 || rank < 0 || rank >= world) { err_ = "pose mailbox: bad rank / world"; return false; }
    name_ = name; rank_ = rank; world_ = world; seq_ = 0;
    if (rank == 0)
    {
        shm_unlink(name.c_str()); // a stale segment of a crashed job
        fd_ = shm_open(name.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
        if (fd_ < 0 || ftruncate(fd_, sizeof(Block)) != 0) { err_ = "pose mailbox: cannot create " + name; close(); return false; }
    }
    else
    {
        fd_ = shm_open(name.c_str(), O_RDWR, 0600);
        if (fd_ < 0) { err_ = "pose mailbox: cannot open " + name; return false; }
    }
    void *p = mmap(nullptr, sizeof(Block), PROT_READ | PROT_WRITE, MAP_SHARED, fd_, 0);
    if (p == MAP_FAILED) { err_ = "pose mailbox: mmap failed"; close(); return false; }
    b_ = static_cast<Block *>(p);
    if (rank == 0) std::memset(p, 0, sizeof(Block)); // ftruncate zero-fills; explicit for clarity
    return true;
}

void PoseMailbox::close()
{
    if (b_) munmap(b_, sizeof(Block));
    b_ = nullptr;
    if (fd_ >= 0) ::close(fd_);
    fd_ = -1;
    if (rank_ == 0 && !name_.empty()) shm_unlink(name_.c_str());
    name_.clear();
}

int PoseMailbox::exchange(float *msg13, double timeout_s)
{
    if (!b_) { err_ = "pose mailbox: not open"; return 1; }
    ++seq_;
    const auto t0 = std::chrono::steady_clock::now();
    unsigned long spins = 0;
    auto late = [&]() {
        if ((++spins & 0xffff) != 0) return false;
        return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > timeout_s;
    };
    if (rank_ == 0)
    {
        // every reader has taken the previous message before it is overwritten
        for (int r = 1; r < world_; ++r)
            while (b_->ack[r].load(std::memory_order_acquire) < seq_ - 1)
            {
                relax();
                if (late()) { err_ = "pose mailbox: a reader never acknowledged"; return 1; }
            }
        std::memcpy(b_->msg, msg13, 13 * sizeof(float));
        b_->seq.store(seq_, std::memory_order_release);
    }
    else
    {
        while (b_->seq.load(std::memory_order_acquire) < seq_)
        {
            relax();
            if (late()) { err_ = "pose mailbox: the tracking rank never published"; return 1; }
        }
        std::memcpy(msg13, b_->msg, 13 * sizeof(float));
        b_->ack[rank_].store(seq_, std::memory_order_release);
    }
    return 0;
}
} // namespace kf
