#!/usr/bin/env python3
"""Derive the compact road-geometry fixture used by the synthetic road-network
scenarios from the reference's CommonRoad map.

Reads  /root/reference/scenarios/road_network/lanelets/offline_road_data/LabMapCommonRoad.xml
Writes p-dmpc_b200/data/lab_map.npz

Run in the build container only (the reference checkout is not on the GPU box);
the .npz (geometry of 104 lanelets + their topology, ~25 KB) is committed.
Parsing mirrors RoadDataCommonRoad.get_lanelets
(scenarios/road_network/lanelets/RoadDataCommonRoad.m:48-66).
"""
import os
import sys
import xml.etree.ElementTree as ET

import numpy as np

SRC = "/root/reference/scenarios/road_network/lanelets/offline_road_data/LabMapCommonRoad.xml"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "p-dmpc_b200", "data", "lab_map.npz")


def main():
    root = ET.parse(SRC).getroot()
    lanelets = root.findall("lanelet")
    n = len(lanelets)
    ids = [int(l.get("id")) for l in lanelets]
    assert ids == list(range(1, n + 1)), "lanelet ids are expected to be 1..n in file order"
    ptr = [0]
    cols = {k: [] for k in ("rx", "ry", "lx", "ly")}
    pred = np.zeros((n, n), dtype=np.uint8)   # pred[i, j] = 1: lanelet j+1 is a predecessor of i+1
    succ = np.zeros((n, n), dtype=np.uint8)
    adj_left = np.zeros(n, dtype=np.int32)    # 0 = none, else 1-based id
    adj_right = np.zeros(n, dtype=np.int32)
    adj_left_same = np.zeros(n, dtype=np.uint8)
    adj_right_same = np.zeros(n, dtype=np.uint8)
    for i, l in enumerate(lanelets):
        def pts(tag):
            b = l.find(tag)
            return [(float(p.find("x").text), float(p.find("y").text)) for p in b.findall("point")]
        lb, rb = pts("leftBound"), pts("rightBound")
        assert len(lb) == len(rb)
        cols["lx"] += [p[0] for p in lb]
        cols["ly"] += [p[1] for p in lb]
        cols["rx"] += [p[0] for p in rb]
        cols["ry"] += [p[1] for p in rb]
        ptr.append(ptr[-1] + len(lb))
        for p in l.findall("predecessor"):
            pred[i, int(p.get("ref")) - 1] = 1
        for s in l.findall("successor"):
            succ[i, int(s.get("ref")) - 1] = 1
        al, ar = l.find("adjacentLeft"), l.find("adjacentRight")
        if al is not None:
            adj_left[i] = int(al.get("ref"))
            adj_left_same[i] = al.get("drivingDir") == "same"
        if ar is not None:
            adj_right[i] = int(ar.get("ref"))
            adj_right_same[i] = ar.get("drivingDir") == "same"
    inter = sorted({int(e.get("ref")) for inc in root.iter("incoming") for e in inc
                    if e.tag in ("incomingLanelet", "successorsRight", "successorsStraight", "successorsLeft")})
    np.savez_compressed(
        DST, ptr=np.asarray(ptr, dtype=np.int32),
        rx=np.asarray(cols["rx"]), ry=np.asarray(cols["ry"]),
        lx=np.asarray(cols["lx"]), ly=np.asarray(cols["ly"]),
        pred=pred, succ=succ, adj_left=adj_left, adj_right=adj_right,
        adj_left_same=adj_left_same, adj_right_same=adj_right_same,
        intersection_lanelets=np.asarray(inter, dtype=np.int32))
    print(f"wrote {os.path.normpath(DST)}: {n} lanelets, {ptr[-1]} points/side, "
          f"{os.path.getsize(DST)} bytes")


if __name__ == "__main__":
    sys.exit(main())
