#include "partition.h"

#include <algorithm>
#include <numeric>

namespace cfb {

Partition Partition::fromOwners(const RoadNet &net, const std::vector<int> &interOwner, int world) {
    Partition p;
    p.world = world;
    p.interOwner = interOwner;
    const int nL = net.nLanes(), nK = net.nLinks();
    p.drvOwner.assign(nL + nK, 0);
    for (int l = 0; l < nL; ++l) p.drvOwner[l] = interOwner[net.roadEndInter[net.laneRoad[l]]];
    for (int k = 0; k < nK; ++k) p.drvOwner[nL + k] = interOwner[net.rlInter[net.llRoadLink[k]]];
    p.boundary.assign(world, std::vector<std::vector<int>>(world));
    for (int l = 0; l < nL; ++l) {
        const int a = interOwner[net.roadStartInter[net.laneRoad[l]]], b = p.drvOwner[l];
        if (a != b) p.boundary[a][b].push_back(l);
    }
    return p;
}

Partition Partition::columnStrips(const RoadNet &net, int world) {
    const int nI = net.nInter();
    std::vector<int> real;
    for (int i = 0; i < nI; ++i)
        if (!net.interVirtual[i]) real.push_back(i);
    std::stable_sort(real.begin(), real.end(), [&](int a, int b) { return net.interPoint[a].x < net.interPoint[b].x; });
    // thresholds: the last x of every chunk of real intersections (chunks never split one x value)
    std::vector<double> limit(world, 0.0);
    const size_t n = real.size();
    size_t pos = 0;
    for (int k = 0; k < world; ++k) {
        size_t end = n * (size_t) (k + 1) / (size_t) world;
        if (end <= pos) end = std::min(n, pos + 1);
        while (end < n && end > 0 && net.interPoint[real[end]].x == net.interPoint[real[end - 1]].x) ++end;
        if (k == world - 1) end = n;
        limit[k] = end > 0 ? net.interPoint[real[std::min(end, n) - 1]].x : 0.0;
        pos = end;
    }
    std::vector<int> owner(nI, world - 1);
    for (int i = 0; i < nI; ++i) {
        for (int k = 0; k < world; ++k)
            if (net.interPoint[i].x <= limit[k]) { owner[i] = k; break; }
    }
    return fromOwners(net, owner, world);
}

int Partition::numBoundaryLanes() const {
    int n = 0;
    for (auto &row : boundary)
        for (auto &v : row) n += (int) v.size();
    return n;
}

std::string Partition::validate(const RoadNet &net, double lookAhead) const {
    for (int a = 0; a < world; ++a)
        for (int b = 0; b < world; ++b)
            for (int l : boundary[a][b]) {
                // A vehicle on a lane of rank a looks at most `lookAhead` metres beyond its lane end;
                // it must never see past the seam lane itself, and a vehicle admitted to the start of
                // the seam lane must never see past the seam lane's end.
                if (net.laneLength[l] <= lookAhead)
                    return "boundary lane " + net.laneName(l) + " is shorter than the leader look-ahead";
            }
    return "";
}

}  // namespace cfb
