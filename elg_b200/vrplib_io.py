"""Minimal CVRPLIB reader: the two calls the reference makes into the un-vendored `vrplib` PyPI package
(`vrplib.read_instance(path)` / `vrplib.read_solution(path)`, CVRP/test_vrplib.py:57,112-113).
Only the fields the reference consumes are produced (CVRP/CVRPEnv.py:87-103): node_coord, demand,
capacity, depot (0-based) and the solution's cost; EUC_2D instances with explicit coordinates."""
import numpy as np


def read_instance(path):
    spec, section = {}, None
    coords, demands, depots = [], [], []
    with open(path) as f:
        for raw in f:
            line = raw.strip()
            if not line or line == "EOF":
                continue
            if line.endswith("_SECTION"):
                section = line
                continue
            if section is None or (":" in line and not line[0].isdigit() and not line[0] == "-"):
                key, _, val = line.partition(":")
                spec[key.strip()] = val.strip().strip('"')
                section = None
                continue
            parts = line.split()
            if section == "NODE_COORD_SECTION":
                coords.append((float(parts[1]), float(parts[2])))
            elif section == "DEMAND_SECTION":
                demands.append(float(parts[1]))
            elif section == "DEPOT_SECTION":
                if int(parts[0]) >= 0:
                    depots.append(int(parts[0]) - 1)
    dim = int(spec.get("DIMENSION", len(coords)))
    if len(coords) != dim or len(demands) != dim:
        raise ValueError("%s: expected %d coordinates/demands, got %d/%d" % (path, dim, len(coords), len(demands)))
    return {"name": spec.get("NAME"), "dimension": dim, "capacity": float(spec["CAPACITY"]),
            "edge_weight_type": spec.get("EDGE_WEIGHT_TYPE"), "node_coord": np.array(coords, dtype=np.float64),
            "demand": np.array(demands, dtype=np.float64), "depot": np.array(depots or [0])}


def read_solution(path):
    routes, cost = [], None
    with open(path) as f:
        for raw in f:
            line = raw.strip()
            if line.lower().startswith("route"):
                routes.append([int(x) for x in line.split(":")[1].split()])
            elif line.lower().startswith("cost"):
                cost = float(line.split()[1])
    return {"routes": routes, "cost": cost}
