// Host-side static road graph: flat tables built from the reference's roadnet JSON format.
//
// Everything the hot path consumes (lane / laneLink lengths, cross ordering and distances,
// phase availability) is derived here with the same FP64 operation order as the reference's
// loader (roadnet.cpp:42-325, :456-505, :515-576; utility.cpp:30-84), so the tables are
// bit-identical to the reference's pointer graph (verified by tests/test_loader.py against
// oracle/_ref/refdump static).  The tables are immutable after load and uploaded once.
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace cfb {

struct Pt {
    double x = 0.0, y = 0.0;
};

enum LinkType : int { TURN_RIGHT = 1, TURN_LEFT = 2, GO_STRAIGHT = 3 };  // roadnet.h:401-403

struct CrossRef {     // one entry of a laneLink's ordered cross list
    int cross;        // global cross id
    int side;         // which side of the cross this laneLink is (0/1)
};

struct RoadNet {
    // ---- roads (file order) ----
    std::vector<std::string> roadId;
    std::vector<int> roadStartInter, roadEndInter;
    std::vector<int> roadLaneBeg;                 // nRoads+1, lanes of road r = [beg[r], beg[r+1])
    std::vector<std::vector<Pt>> roadPoints;
    std::map<std::string, int> roadIndex;

    // ---- intersections (file order) ----
    std::vector<std::string> interId;
    std::vector<uint8_t> interVirtual;
    std::vector<double> interWidth;
    std::vector<Pt> interPoint;
    std::vector<std::vector<int>> interRoads;     // roads listed under "roads"
    std::vector<int> interRoadLinkBeg;            // nInter+1 (global roadLink ids, contiguous)
    std::vector<int> interLinkBeg;                // nInter+1 (laneLink ids 0-based within links)
    std::vector<int> interCrossBeg;               // nInter+1
    std::vector<int> interPhaseBeg;               // nInter+1 into phaseTime
    std::map<std::string, int> interIndex;

    // ---- roadLinks (global id = intersection order x index) ----
    std::vector<int> rlType, rlStartRoad, rlEndRoad, rlInter, rlLinkBeg;  // rlLinkBeg: nRL+1

    // ---- traffic-light phases ----
    std::vector<double> phaseTime;                // per global phase
    std::vector<int> phaseAvailBeg;               // per global phase: offset into phaseAvail
    std::vector<uint8_t> phaseAvail;              // [phase][roadLink-in-intersection]

    // ---- lanes (global id = roads in file order x lane index; == roadnet.getLanes()) ----
    std::vector<int> laneRoad, laneIdx;
    std::vector<double> laneWidth, laneMaxSpeed, laneLength;
    std::vector<std::vector<Pt>> lanePoints;
    std::vector<std::vector<int>> laneOutLinks;   // laneLink ids (0-based) in push order (roadnet.cpp:253)

    // ---- laneLinks (global id = intersections x roadLinks x laneLinks; == getLaneLinks()) ----
    std::vector<int> llStartLane, llEndLane, llRoadLink;
    std::vector<double> llLength;
    std::vector<std::vector<Pt>> llPoints;
    std::vector<std::vector<CrossRef>> llCrosses; // ascending distance along this link

    // ---- crosses (global id = intersections x creation order) ----
    std::vector<int> crossLink[2];                // laneLink ids
    std::vector<double> crossDist[2];             // distanceOnLane

    int nRoads() const { return (int) roadId.size(); }
    int nInter() const { return (int) interId.size(); }
    int nLanes() const { return (int) laneRoad.size(); }
    int nLinks() const { return (int) llStartLane.size(); }
    int nRoadLinks() const { return (int) rlType.size(); }
    int nCross() const { return (int) crossLink[0].size(); }
    int nDrivables() const { return nLanes() + nLinks(); }

    std::string laneName(int lane) const { return roadId[laneRoad[lane]] + "_" + std::to_string(laneIdx[lane]); }
    int laneOf(int road, int idx) const { return roadLaneBeg[road] + idx; }
    int roadNumLanes(int road) const { return roadLaneBeg[road + 1] - roadLaneBeg[road]; }
    bool linkIsTurn(int ll) const { int t = rlType[llRoadLink[ll]]; return t == TURN_LEFT || t == TURN_RIGHT; }
    // laneLinks of `lane` whose end lane belongs to `road` (Lane::getLaneLinksToRoad, roadnet.cpp:447)
    void linksToRoad(int lane, int road, std::vector<int> &out) const {
        out.clear();
        for (int ll : laneOutLinks[lane])
            if (laneRoad[llEndLane[ll]] == road) out.push_back(ll);
    }
    bool roadConnected(int a, int b) const {  // Road::connectedToRoad, roadnet.cpp:736
        for (int l = roadLaneBeg[a]; l < roadLaneBeg[a + 1]; ++l)
            for (int ll : laneOutLinks[l])
                if (laneRoad[llEndLane[ll]] == b) return true;
        return false;
    }
    double roadAverageLength(int r) const {  // Road::averageLength, roadnet.cpp:709
        double sum = 0;
        int n = roadNumLanes(r);
        if (n == 0) return 0;
        for (int l = roadLaneBeg[r]; l < roadLaneBeg[r + 1]; ++l) sum += laneLength[l];
        return sum / n;
    }

    // Returns false (message on stderr) on a format error, like RoadNet::loadFromJson.
    bool load(const std::string &path);
};

}  // namespace cfb
