// pose_mailbox.hpp -- {tracking_ok, 4x3 pose} from the tracking rank to the slab ranks of one node (SURVEY.md §8e:
// "ICP stays on one GPU, and only the 4x4 pose is broadcast").  52 bytes through POSIX shared memory: rank 0
// publishes payload then sequence number, every other rank spins on the sequence number and acknowledges, so that the
// payload is never overwritten before everybody has taken it.  This is the native kf::ShardComm::broadcast_pose of a
// sharded kf::kinectfusion (all ranks of a z-slab sharded volume sit on one NVLink box); a launcher that spans nodes
// supplies its own callback instead.  No reference counterpart: the reference is single-GPU.
#pragma once
#include <atomic>
#include <string>

namespace kf
{
class PoseMailbox
{
    struct Block // one cache line per writer
    {
        alignas(64) std::atomic<long long> seq;
        alignas(64) float msg[16];
        alignas(64) std::atomic<long long> ack[64];
    };
    Block *b_ = nullptr;
    int fd_ = -1, rank_ = 0, world_ = 1;
    long long seq_ = 0;
    std::string name_, err_;

public:
    PoseMailbox() = default;
    PoseMailbox(const PoseMailbox &) = delete;
    PoseMailbox &operator=(const PoseMailbox &) = delete;
    ~PoseMailbox() { close(); }

    // Rank 0 creates (and zeroes) the segment; call it on rank 0 first, then -- after a barrier of the launcher's --
    // on the other ranks.  name: "/something" unique per job.  world <= 64.
    bool open(const std::string &name, int rank, int world);
    void close();
    // msg13 valid on rank 0 on entry, on every rank on return.  0 ok, 1 = a peer did not arrive within timeout_s.
    int exchange(float *msg13, double timeout_s = 20.0);
    const std::string &lastError() const { return err_; }

    // kf::ShardComm::broadcast_pose with user = PoseMailbox*
    static int callback(float *msg13, void *user) { return static_cast<PoseMailbox *>(user)->exchange(msg13); }
};
} // namespace kf
