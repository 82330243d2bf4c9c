// Builds the flat static tables of roadnet.h from a roadnet JSON file.
// Geometry follows the reference's arithmetic order exactly (see roadnet.h header comment).
#include "roadnet.h"

#include <algorithm>
#include <cmath>
#include <iostream>

#include "json_min.h"

namespace cfb {
namespace {

// ---- 2-D helpers (utility.cpp:14-84; FP64 operation order preserved) ----
inline Pt sub(Pt a, Pt b) { return {a.x - b.x, a.y - b.y}; }
inline Pt add(Pt a, Pt b) { return {a.x + b.x, a.y + b.y}; }
inline Pt mul(Pt a, double k) { return {a.x * k, a.y * k}; }
inline double norm(Pt a) { return sqrt(a.x * a.x + a.y * a.y); }
inline Pt unit(Pt a) { double l = norm(a); return {a.x / l, a.y / l}; }
inline Pt negNormal(Pt a) { return {-(-a.y), -(a.x)}; }  // -(u.normal())
inline double cross2(Pt a, Pt b) { return a.x * b.y - a.y * b.x; }
inline double dot2(Pt a, Pt b) { return a.x * b.x + a.y * b.y; }
constexpr double kEps = 1e-8;
inline int sgn(double x) { return (x + kEps > 0) - (x < kEps); }
inline double max2(double x, double y) { return x > y ? x : y; }
inline double min2(double x, double y) { return x < y ? x : y; }

double polyLength(const std::vector<Pt> &p) {  // roadnet.cpp:31-36
    double length = 0.0;
    for (size_t i = 0; i + 1 < p.size(); i++) length += norm(sub(p[i + 1], p[i]));
    return length;
}

Pt pointAt(const std::vector<Pt> &p, double dis) {  // roadnet.cpp:17-29
    dis = min2(max2(dis, 0), polyLength(p));
    if (dis <= 0.0) return p[0];
    for (size_t i = 1; i < p.size(); i++) {
        double len = norm(sub(p[i - 1], p[i]));
        if (dis > len) dis -= len;
        else return add(p[i - 1], mul(sub(p[i], p[i - 1]), dis / len));
    }
    return p.back();
}

Pt directionAt(const std::vector<Pt> &p, double dis) {  // roadnet.cpp:400-410
    double remain = dis;
    for (int i = 0; i + 1 < (int) p.size(); i++) {
        double len = norm(sub(p[i + 1], p[i]));
        if (remain < len) return unit(sub(p[i + 1], p[i]));
        remain -= len;
    }
    return unit(sub(p[p.size() - 1], p[p.size() - 2]));
}

Pt lerp(Pt p1, Pt p2, double a) { return {(p2.x - p1.x) * a + p1.x, (p2.y - p1.y) * a + p1.y}; }

struct FormatError : std::runtime_error {
    explicit FormatError(const std::string &m) : std::runtime_error(m) {}
};

const Json &member(const Json &o, const char *name) {
    const Json *v = o.find(name);
    if (!v) throw FormatError(std::string(name) + " is required but missing in json file");
    return *v;
}
const Json &memberArray(const Json &o, const char *name) {
    const Json &v = member(o, name);
    if (!v.isArray()) throw FormatError(std::string(name) + ": expected type array");
    return v;
}
const Json &memberObject(const Json &o, const char *name) {
    const Json &v = member(o, name);
    if (!v.isObject()) throw FormatError(std::string(name) + ": expected type object");
    return v;
}
double memberDouble(const Json &o, const char *name) {
    const Json &v = member(o, name);
    if (!v.isNumber()) throw FormatError(std::string(name) + ": expected type d");
    return v.asDouble();
}
int memberInt(const Json &o, const char *name) {
    const Json &v = member(o, name);
    if (!v.isInt()) throw FormatError(std::string(name) + ": expected type i");
    return v.asInt();
}
bool memberBool(const Json &o, const char *name) {
    const Json &v = member(o, name);
    if (!v.isBool()) throw FormatError(std::string(name) + ": expected type b");
    return v.asBool();
}
const std::string &memberString(const Json &o, const char *name) {
    const Json &v = member(o, name);
    if (!v.isString()) throw FormatError(std::string(name) + ": expected type PKc");
    return v.s;
}

}  // namespace

// Lane centre-lines of one road, trimmed by the (currently known) intersection widths
// (Road::initLanesPoints, roadnet.cpp:456-505).
static void buildLaneGeometry(RoadNet &n, int r) {
    std::vector<Pt> rp = n.roadPoints[r];
    if (!n.interVirtual[n.roadStartInter[r]]) {
        double w = n.interWidth[n.roadStartInter[r]];
        Pt p1 = rp[0], p2 = rp[1];
        rp[0] = add(p1, mul(unit(sub(p2, p1)), w));
    }
    if (!n.interVirtual[n.roadEndInter[r]]) {
        double w = n.interWidth[n.roadEndInter[r]];
        Pt p1 = rp[rp.size() - 2], p2 = rp[rp.size() - 1];
        rp[rp.size() - 1] = sub(p2, mul(unit(sub(p2, p1)), w));
    }
    double dsum = 0.0;
    for (int l = n.roadLaneBeg[r]; l < n.roadLaneBeg[r + 1]; ++l) {
        double dmin = dsum, dmax = dsum + n.laneWidth[l];
        std::vector<Pt> &lp = n.lanePoints[l];
        lp.clear();
        const int m = (int) rp.size();
        for (int j = 0; j < m; ++j) {
            Pt u;
            if (j == 0) u = unit(sub(rp[1], rp[0]));
            else if (j + 1 == m) u = unit(sub(rp[j], rp[j - 1]));
            else {
                Pt u1 = unit(sub(rp[j + 1], rp[j]));
                Pt u2 = unit(sub(rp[j], rp[j - 1]));
                u = unit(add(u1, u2));
            }
            Pt v = negNormal(u);
            lp.push_back(add(rp[j], mul(v, (dmin + dmax) / 2.0)));
        }
        n.laneLength[l] = polyLength(lp);
        dsum += n.laneWidth[l];
    }
}

// Pairwise first intersection of laneLink polylines inside one intersection
// (Intersection::initCrosses, roadnet.cpp:515-576).
static void buildCrosses(RoadNet &n, int inter) {
    const int lb0 = n.interLinkBeg[inter], lb1 = n.interLinkBeg[inter + 1];
    for (int a = lb0; a < lb1; ++a) {
        for (int b = a + 1; b < lb1; ++b) {
            const std::vector<Pt> &va = n.llPoints[a], &vb = n.llPoints[b];
            double disa = 0.0;
            bool found = false;
            for (int ia = 0; ia + 1 < (int) va.size() && !found; ia++) {
                double disb = 0.0;
                for (int ib = 0; ib + 1 < (int) vb.size(); ib++) {
                    Pt A1 = va[ia], A2 = va[ia + 1], B1 = vb[ib], B2 = vb[ib + 1];
                    // parallel pieces are skipped *without* advancing disb (as the reference does)
                    if (sgn(cross2(sub(A2, A1), sub(B2, B1))) == 0) continue;
                    Pt u = sub(A2, A1), v = sub(B2, B1);
                    Pt P = add(A1, mul(u, cross2(sub(B1, A1), v) / cross2(u, v)));
                    bool onA = sgn(cross2(sub(A2, A1), sub(P, A1))) == 0 && sgn(dot2(sub(P, A1), sub(P, A2))) <= 0;
                    bool onB = sgn(cross2(sub(B2, B1), sub(P, B1))) == 0 && sgn(dot2(sub(P, B1), sub(P, B2))) <= 0;
                    if (onA && onB) {
                        n.crossLink[0].push_back(a);
                        n.crossLink[1].push_back(b);
                        n.crossDist[0].push_back(disa + norm(sub(P, A1)));
                        n.crossDist[1].push_back(disb + norm(sub(P, B1)));
                        found = true;
                        break;
                    }
                    disb += norm(sub(vb[ib + 1], vb[ib]));
                }
                if (!found) disa += norm(sub(va[ia + 1], va[ia]));
            }
        }
    }
    const int c0 = n.interCrossBeg[inter], c1 = (int) n.crossLink[0].size();
    for (int c = c0; c < c1; ++c) {
        n.llCrosses[n.crossLink[0][c]].push_back({c, 0});
        n.llCrosses[n.crossLink[1][c]].push_back({c, 1});
    }
    for (int l = lb0; l < lb1; ++l) {
        auto &v = n.llCrosses[l];
        // same algorithm (std::sort) on the same input order as the reference => same tie order
        std::sort(v.begin(), v.end(), [&n](const CrossRef &ca, const CrossRef &cb) {
            return n.crossDist[ca.side][ca.cross] < n.crossDist[cb.side][cb.cross];
        });
    }
}

bool RoadNet::load(const std::string &path) {
    bool opened = false;
    Json doc = Json::parseFile(path, &opened);  // parse errors propagate (JsonFormatError in the reference)
    if (!opened) {
        std::cerr << "cannot open roadnet file" << std::endl;
        return false;
    }
    if (!doc.isObject()) throw JsonError("roadnet config file: expected type object");
    std::string where;
    try {
        const Json &interV = memberArray(doc, "intersections");
        const Json &roadV = memberArray(doc, "roads");
        const int nR = (int) roadV.arr.size(), nI = (int) interV.arr.size();
        roadId.resize(nR);
        roadStartInter.assign(nR, -1);
        roadEndInter.assign(nR, -1);
        roadPoints.resize(nR);
        interId.resize(nI);
        interVirtual.assign(nI, 0);
        interWidth.assign(nI, 0.0);
        interPoint.resize(nI);
        interRoads.resize(nI);
        for (int i = 0; i < nR; ++i) {
            where = "road[" + std::to_string(i) + "]";
            roadId[i] = memberString(roadV.arr[i], "id");
            roadIndex[roadId[i]] = i;
        }
        for (int i = 0; i < nI; ++i) {
            where = "intersection[" + std::to_string(i) + "]";
            interId[i] = memberString(interV.arr[i], "id");
            interIndex[interId[i]] = i;
        }
        // ---- roads, lanes, road points ----
        roadLaneBeg.assign(1, 0);
        for (int i = 0; i < nR; ++i) {
            where = "roads/" + roadId[i];
            const Json &rv = roadV.arr[i];
            if (!rv.isObject()) throw FormatError("road[" + std::to_string(i) + "]: expected type object");
            auto si = interIndex.find(memberString(rv, "startIntersection"));
            auto ei = interIndex.find(memberString(rv, "endIntersection"));
            if (si == interIndex.end()) throw FormatError("startIntersection does not exist.");
            if (ei == interIndex.end()) throw FormatError("endIntersection does not exist.");
            roadStartInter[i] = si->second;
            roadEndInter[i] = ei->second;
            int li = 0;
            for (const Json &lv : memberArray(rv, "lanes").arr) {
                if (!lv.isObject()) throw FormatError("lane: expected type object");
                laneWidth.push_back(memberDouble(lv, "width"));
                laneMaxSpeed.push_back(memberDouble(lv, "maxSpeed"));
                laneRoad.push_back(i);
                laneIdx.push_back(li++);
            }
            roadLaneBeg.push_back((int) laneRoad.size());
            for (const Json &pv : memberArray(rv, "points").arr) {
                if (!pv.isObject()) throw FormatError("point of road: expected type object");
                Pt p;
                p.x = memberDouble(pv, "x");
                p.y = memberDouble(pv, "y");
                roadPoints[i].push_back(p);
            }
            if (roadPoints[i].size() < 2) throw FormatError("road needs at least 2 points");
        }
        const int nL = (int) laneRoad.size();
        laneLength.assign(nL, 0.0);
        lanePoints.resize(nL);
        laneOutLinks.resize(nL);
        // first geometry pass: intersection widths are not known yet (all zero, nothing virtual),
        // exactly the state the reference is in at roadnet.cpp:127-129
        for (int i = 0; i < nR; ++i) buildLaneGeometry(*this, i);

        // ---- intersections, roadLinks, laneLinks, lights ----
        interRoadLinkBeg.assign(1, 0);
        interLinkBeg.assign(1, 0);
        interPhaseBeg.assign(1, 0);
        rlLinkBeg.assign(1, 0);
        for (int i = 0; i < nI; ++i) {
            where = "intersections/" + interId[i];
            const Json &iv = interV.arr[i];
            if (!iv.isObject()) throw FormatError("intersection: expected type object");
            const Json &pv = memberObject(iv, "point");
            interVirtual[i] = memberBool(iv, "virtual") ? 1 : 0;
            interPoint[i].x = memberDouble(pv, "x");
            interPoint[i].y = memberDouble(pv, "y");
            for (const Json &rn : memberArray(iv, "roads").arr) {
                auto it = roadIndex.find(rn.s);
                if (it == roadIndex.end()) throw FormatError("No such road: " + rn.s);
                interRoads[i].push_back(it->second);
            }
            if (!interVirtual[i]) {
                interWidth[i] = memberDouble(iv, "width");
                int rlIdx = 0;
                for (const Json &rlv : memberArray(iv, "roadLinks").arr) {
                    where = "intersections/" + interId[i] + "/roadLinks[" + std::to_string(rlIdx) + "]";
                    if (!rlv.isObject()) throw FormatError("roadLink: expected type object");
                    const std::string &ty = memberString(rlv, "type");
                    int t;
                    if (ty == "turn_left") t = TURN_LEFT;
                    else if (ty == "turn_right") t = TURN_RIGHT;
                    else if (ty == "go_straight") t = GO_STRAIGHT;
                    else throw FormatError("unknown roadLink type: " + ty);
                    auto sr = roadIndex.find(memberString(rlv, "startRoad"));
                    auto er = roadIndex.find(memberString(rlv, "endRoad"));
                    if (sr == roadIndex.end() || er == roadIndex.end()) throw FormatError("No such road in roadLink");
                    const int rl = (int) rlType.size();
                    rlType.push_back(t);
                    rlStartRoad.push_back(sr->second);
                    rlEndRoad.push_back(er->second);
                    rlInter.push_back(i);
                    for (const Json &llv : memberArray(rlv, "laneLinks").arr) {
                        if (!llv.isObject()) throw FormatError("laneLink: expected type object");
                        int sIdx = memberInt(llv, "startLaneIndex");
                        int eIdx = memberInt(llv, "endLaneIndex");
                        if (sIdx < 0 || sIdx >= roadNumLanes(sr->second)) throw FormatError("startLaneIndex out of range");
                        if (eIdx < 0 || eIdx >= roadNumLanes(er->second)) throw FormatError("startLaneIndex out of range");
                        const int sl = laneOf(sr->second, sIdx), el = laneOf(er->second, eIdx);
                        std::vector<Pt> pts;
                        const Json *pj = llv.find("points");
                        if (pj && !pj->isArray()) throw FormatError("points in laneLink: expected type array");
                        if (pj && !pj->arr.empty()) {
                            for (const Json &q : pj->arr) {
                                Pt p;
                                p.x = memberDouble(q, "x");
                                p.y = memberDouble(q, "y");
                                pts.push_back(p);
                            }
                        } else {
                            // default curve: a 3-segment control polygon smoothed by repeated
                            // interpolation (roadnet.cpp:212-247); uses first-pass lane geometry
                            const double w = interWidth[i];
                            Pt start = pointAt(lanePoints[sl], laneLength[sl] - w);
                            Pt end = pointAt(lanePoints[el], 0.0 + w);
                            double len = norm(Pt{end.x - start.x, end.y - start.y});
                            Pt sd = directionAt(lanePoints[sl], laneLength[sl] - w);
                            Pt ed = directionAt(lanePoints[el], 0.0 + w);
                            double minGap = 5;
                            double g1x = sd.x * len * 0.5, g1y = sd.y * len * 0.5;
                            double g2x = -ed.x * len * 0.5, g2y = -ed.y * len * 0.5;
                            if (g1x * g1x + g1y * g1y < 25 && w >= 5) { g1x = minGap * sd.x; g1y = minGap * sd.y; }
                            if (g2x * g2x + g2y * g2y < 25 && w >= 5) { g2x = minGap * ed.x; g2y = minGap * ed.y; }
                            Pt mid1{start.x + g1x, start.y + g1y}, mid2{end.x + g2x, end.y + g2y};
                            const int np = 10;
                            for (int k = 0; k <= np; k++) {
                                double a = k / double(np);
                                Pt p1 = lerp(start, mid1, a), p2 = lerp(mid1, mid2, a), p3 = lerp(mid2, end, a);
                                Pt p4 = lerp(p1, p2, a), p5 = lerp(p2, p3, a);
                                pts.push_back(lerp(p4, p5, a));
                            }
                        }
                        const int ll = (int) llStartLane.size();
                        llStartLane.push_back(sl);
                        llEndLane.push_back(el);
                        llRoadLink.push_back(rl);
                        llLength.push_back(polyLength(pts));
                        llPoints.push_back(std::move(pts));
                        laneOutLinks[sl].push_back(ll);
                    }
                    rlLinkBeg.push_back((int) llStartLane.size());
                    ++rlIdx;
                }
                where = "intersections/" + interId[i] + "/trafficLight";
                const Json &tl = memberObject(iv, "trafficLight");
                const int nRL = (int) rlType.size() - interRoadLinkBeg[i];
                for (const Json &ph : memberArray(tl, "lightphases").arr) {
                    if (!ph.isObject()) throw FormatError("lightphase: expected type object");
                    phaseTime.push_back(memberDouble(ph, "time"));
                    phaseAvailBeg.push_back((int) phaseAvail.size());
                    phaseAvail.resize(phaseAvail.size() + nRL, 0);
                    for (const Json &av : memberArray(ph, "availableRoadLinks").arr) {
                        if (!av.isInt()) throw FormatError("availableRoadLink: expected type int");
                        size_t k = av.asUint();
                        if (k >= (size_t) nRL) throw FormatError("index out of range");
                        phaseAvail[phaseAvailBeg.back() + k] = 1;
                    }
                }
                if (phaseTime.size() == (size_t) interPhaseBeg[i])
                    throw FormatError("non-virtual intersection needs at least one light phase");
            }
            interRoadLinkBeg.push_back((int) rlType.size());
            interLinkBeg.push_back((int) llStartLane.size());
            interPhaseBeg.push_back((int) phaseTime.size());
        }
    } catch (const FormatError &e) {
        std::cerr << "Error occurred when reading the roadnet file: " << std::endl;
        std::cerr << "/" << where << " " << e.what() << std::endl;
        return false;
    }
    llCrosses.resize(llStartLane.size());
    interCrossBeg.assign(1, 0);
    for (int i = 0; i < nInter(); ++i) {
        buildCrosses(*this, i);
        interCrossBeg.push_back((int) crossLink[0].size());
    }
    // second geometry pass with the real widths / virtual flags (roadnet.cpp:307-308)
    for (int i = 0; i < nRoads(); ++i) buildLaneGeometry(*this, i);
    return true;
}

}  // namespace cfb
