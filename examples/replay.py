"""saveReplay: the engine writes the reference's two replay files (roadnetLogFile once,
replayLogFile one line per step) for the reference's frontend viewer."""
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cityflow  # noqa: E402
from cityflow_b200 import scenario  # noqa: E402

with tempfile.TemporaryDirectory() as d:
    net = scenario.grid_roadnet(3, 3)
    flows = scenario.random_walk_flows(net, frac=1.0, interval=5.0, seed=1)
    cfg = scenario.write_scenario(d, net, flows, save_replay=True)
    c = json.load(open(cfg))
    eng = cityflow.Engine(cfg)
    for t in range(100):
        eng.next_step()
    eng.set_save_replay(False)          # pause logging
    for t in range(50):
        eng.next_step()
    eng.set_save_replay(True)
    eng.set_replay_file("replay_part2.txt")   # relative to the config's "dir"
    for t in range(50):
        eng.next_step()
    del eng
    static = json.load(open(c["dir"] + c["roadnetLogFile"]))["static"]
    lines = open(c["dir"] + c["replayLogFile"]).read().splitlines()
    print("roadnet log: %d nodes, %d edges; replay: %d + %d lines" % (
        len(static["nodes"]), len(static["edges"]), len(lines), len(open(c["dir"] + "replay_part2.txt").read().splitlines())))
    print("last line, first vehicle:", lines[-1].split(",")[0])
