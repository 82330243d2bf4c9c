// Static cut of the road graph across GPUs: every intersection belongs to one
// rank; a laneLink / cross / traffic light follows its intersection; a lane follows its END
// intersection (its head interacts with that intersection's links, crosses and light).
//
// A lane L whose START intersection belongs to rank A and whose END intersection to rank B != A
// is a boundary lane of the ordered pair (A -> B): A "feeds" it (all laneLinks that end in L sit
// in A's intersection), B owns it.  What crosses the seam, once per phase boundary:
//   A -> B  the vehicles leaving A's laneLinks into L during the step (mover records)
//   B -> A  L's tail record + the tail vehicle's fields (Lane::canEnter, leader search and
//           Cross::notify source 1 on A read exactly that vehicle)
// plus the network-wide list of blocker changes (Cross::canPass walks blocker chains that may
// leave the rank).  The reference has no counterpart: its threads share one address space.
#pragma once
#include <string>
#include <vector>

#include "roadnet.h"

namespace cfb {

struct Partition {
    int world = 1;
    std::vector<int> interOwner;                       // per intersection
    std::vector<int> drvOwner;                         // per drivable (lanes first, then laneLinks)
    // boundary[a][b]: lanes fed by rank a and owned by rank b (a != b), ascending lane id
    std::vector<std::vector<std::vector<int>>> boundary;

    // Contiguous strips along x (columns of a grid), balanced by the number of real intersections.
    static Partition columnStrips(const RoadNet &net, int world);
    static Partition fromOwners(const RoadNet &net, const std::vector<int> &interOwner, int world);
    // Empty string if the cut is usable; otherwise what is wrong (e.g. a seam lane shorter than the
    // leader look-ahead, which would need second-order ghost data).
    std::string validate(const RoadNet &net, double lookAhead) const;
    int numBoundaryLanes() const;
};

// Where rank `me`'s seam messages land in the receivers' mailboxes (device_shard.cuh).  A rank lists the lanes it feeds
// ("out", grouped by owner, ascending lane id inside a group) and the lanes it owns that a peer feeds ("in", grouped by
// feeder); a mover message for out-entry j goes to entry outDst[j] of rank outPeer[j]'s IN list, a tail message for
// in-entry j to entry inDst[j] of rank inPeer[j]'s OUT list.  Pure function of the boundary-size matrix
// bsize[a][b] = |boundary[a][b]|, which every rank derives identically -- tests/test_dist_cpu.py lets two processes
// compute their tables independently and checks that they meet.
struct SeamTables {
    std::vector<int> nbr;                         // ranks sharing a seam lane with `me`
    std::vector<int> outPeer, outDst, inPeer, inDst;
};
inline SeamTables seamTables(const std::vector<std::vector<int>> &bsize, int me) {
    SeamTables t;
    const int W = (int) bsize.size();
    for (int q = 0; q < W; ++q) {
        if (q == me) continue;
        if (bsize[me][q] + bsize[q][me] > 0) t.nbr.push_back(q);
        int inBegAtQ = 0, outBegAtQ = 0;   // q's inBeg[me], q's outBeg[me]
        for (int p = 0; p < me; ++p) { inBegAtQ += bsize[p][q]; outBegAtQ += bsize[q][p]; }
        for (int k = 0; k < bsize[me][q]; ++k) { t.outPeer.push_back(q); t.outDst.push_back(inBegAtQ + k); }
        for (int k = 0; k < bsize[q][me]; ++k) { t.inPeer.push_back(q); t.inDst.push_back(outBegAtQ + k); }
    }
    return t;
}

}  // namespace cfb
