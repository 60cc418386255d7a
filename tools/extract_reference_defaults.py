#!/usr/bin/env python
"""Extracts, from the REFERENCE's own sources, every default the drop-in boundary has to reproduce, and writes them to
tests/golden/reference_defaults.json — a fixture pinned on the reference itself (not on the oracle):

  * the default member initialisers of ct_icp::CTICPOptions (include/ct_icp/ct_icp.h), ct_icp::OdometryOptions
    (include/ct_icp/odometry.h), MultipleResolutionVoxelMap::Options + ResolutionParam (include/ct_icp/map.h),
    PreviousFrameMotionModel::Options (include/ct_icp/motion_model.h), the two neighborhood-strategy option structs
    (include/ct_icp/neighborhood_strategy.h) and the enumerators of the enums those fields use;
  * the assignments of the three profile functions (src/ct_icp/odometry.cpp: DefaultDrivingProfile,
    RobustDrivingProfile, DefaultRobustOutdoorLowInertia), in source order (a later assignment overrides an earlier one);
  * the odometry / ct_icp sections of the shipped configurations config/odometry/driving_config.yaml and nclt_config.yaml.

usage: python tools/extract_reference_defaults.py [/root/reference] [out.json]
tests/test_reference_defaults.py compares the engine's cticp_default_* / cticp_profile_* with this file, and — when the
reference tree is present — checks that the file is what this script extracts today."""
import json
import os
import re
import sys


def strip_comments(txt):
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return re.sub(r"//[^\n]*", "", txt)


def struct_body(txt, header_regex):
    m = re.search(header_regex, txt)
    if not m:
        raise RuntimeError("struct not found: " + header_regex)
    i = txt.index("{", m.end() - 1)
    depth, j = 0, i
    while True:
        if txt[j] == "{":
            depth += 1
        elif txt[j] == "}":
            depth -= 1
            if depth == 0:
                return txt[i + 1:j]
        j += 1


def literal(v):
    v = v.strip()
    if v in ("true", "false"):
        return v == "true"
    if re.fullmatch(r"[-+]?\d+", v):
        return int(v)
    if re.fullmatch(r"[-+]?(\d+\.\d*|\.\d+|\d+)([eE][-+]?\d+)?f?", v):
        return float(v.rstrip("f"))
    return v   # an enumerator or an expression: kept as text


def member_defaults(body):
    """`type name = value;` at depth 0 of a struct body (nested structs / functions skipped)."""
    out, depth, stmt = {}, 0, ""
    for ch in body:
        if ch == "{":
            depth += 1
        elif ch == "}":
            depth -= 1
            stmt = ""
        elif depth == 0:
            if ch == ";":
                m = re.fullmatch(r"\s*(?:static\s+|const\s+)*[\w:<>\s]+?[\s&*]+(\w+)\s*=\s*([^=].*?)\s*", stmt, flags=re.S)
                if m and "(" not in m.group(2):
                    out[m.group(1)] = literal(m.group(2))
                stmt = ""
            else:
                stmt += ch
    return out


def enum_values(txt, name):
    body = struct_body(txt, r"enum\s+(?:class\s+)?%s\s*\{" % name)
    names, val, out = [x.strip() for x in body.split(",") if x.strip()], 0, {}
    for n in names:
        if "=" in n:
            n, v = [x.strip() for x in n.split("=")]
            if not re.fullmatch(r"\d+", v):
                continue   # (flag combinations: not option values)
            val = int(v)
        out[n] = val
        val += 1
    return out


def profile_assignments(cpp, fn):
    body = struct_body(cpp, r"OdometryOptions\s+OdometryOptions::%s\s*\(\s*\)\s*\{" % fn)
    out = []
    for m in re.finditer(r"(?:default_options|options)\.([\w.]+)\s*=\s*([^;]+);|ct_icp_options\.(\w+)\s*=\s*([^;]+);", body):
        if m.group(1):
            out.append([m.group(1), literal(m.group(2))])
        else:
            out.append(["ct_icp_options." + m.group(3), literal(m.group(4))])
    return out


def main(ref, out_path):
    rd = lambda p: strip_comments(open(os.path.join(ref, p)).read())
    icp_h, odo_h, map_h = rd("include/ct_icp/ct_icp.h"), rd("include/ct_icp/odometry.h"), rd("include/ct_icp/map.h")
    mm_h, ns_h, odo_cpp = rd("include/ct_icp/motion_model.h"), rd("include/ct_icp/neighborhood_strategy.h"), rd("src/ct_icp/odometry.cpp")
    cost_h = rd("include/ct_icp/cost_functions.h")
    out = {
        "source": "jedeschaud/ct_icp (reference tree): include/ct_icp/{ct_icp,cost_functions,odometry,map,motion_model,neighborhood_strategy}.h, "
                  "src/ct_icp/odometry.cpp, config/odometry/*.yaml — extracted by tools/extract_reference_defaults.py",
        "CTICPOptions": member_defaults(struct_body(icp_h, r"struct\s+CTICPOptions\s*\{")),
        "OdometryOptions": member_defaults(struct_body(odo_h, r"struct\s+OdometryOptions\s*\{")),
        "MapOptions": member_defaults(struct_body(map_h, r"struct\s+Options\s*:\s*public\s+IMapOptions\s*\{")),
        "ResolutionParam": member_defaults(struct_body(map_h, r"struct\s+ResolutionParam\s*\{")),
        "MotionModelOptions": member_defaults(struct_body(mm_h, r"struct\s+Options\s*\{")),
        "enums": {
            "CT_ICP_SOLVER": enum_values(icp_h, "CT_ICP_SOLVER"), "LEAST_SQUARES": enum_values(icp_h, "LEAST_SQUARES"),
            "WEIGHTING_SCHEME": enum_values(icp_h, "WEIGHTING_SCHEME"), "POSE_PARAMETRIZATION": enum_values(cost_h, "POSE_PARAMETRIZATION"),
            "ICP_DISTANCE": enum_values(cost_h, "ICP_DISTANCE"), "MOTION_COMPENSATION": enum_values(odo_h, "MOTION_COMPENSATION"),
            "INITIALIZATION": enum_values(odo_h, "INITIALIZATION"), "SAMPLING_OPTION": enum_values(odo_h, "SAMPLING_OPTION"),
            "MODEL_TYPE": enum_values(mm_h, "MODEL_TYPE"),
        },
        "profiles": {fn: profile_assignments(odo_cpp, fn)
                     for fn in ("DefaultDrivingProfile", "RobustDrivingProfile", "DefaultRobustOutdoorLowInertia")},
    }
    # resolutions = { ResolutionParam{0.2, 0.03, 50}, ... }
    body = struct_body(map_h, r"struct\s+Options\s*:\s*public\s+IMapOptions\s*\{")
    out["MapOptions"]["resolutions"] = [[literal(x) for x in m.group(1).split(",")]
                                        for m in re.finditer(r"ResolutionParam\s*\{([^}]*)\}", body)]
    # neighborhood strategies: the option structs of the two strategies
    out["NeighborStrategyOptions"] = {}
    for m in re.finditer(r"struct\s+Options\s*:\s*(?:public\s+)?INeighborStrategyOptions\s*\{", ns_h):
        b = struct_body(ns_h[m.start():], r"struct\s+Options\s*:\s*(?:public\s+)?INeighborStrategyOptions\s*\{")
        d = member_defaults(b)
        t = re.search(r'return\s+"(\w+)"', b)
        out["NeighborStrategyOptions"][t.group(1) if t else "strategy_%d" % len(out["NeighborStrategyOptions"])] = d
    out["NeighborStrategyOptions"]["base"] = member_defaults(struct_body(ns_h, r"struct\s+INeighborStrategyOptions\s*\{"))
    import yaml
    out["yaml"] = {}
    for name in ("driving_config", "nclt_config"):
        y = yaml.safe_load(open(os.path.join(ref, "config/odometry/%s.yaml" % name)))
        out["yaml"][name] = y.get("odometry_options", y)
    with open(out_path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
        f.write("\n")
    return out


if __name__ == "__main__":
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    dst = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                            "tests", "golden", "reference_defaults.json")
    o = main(ref, dst)
    print("wrote", dst, {k: (len(v) if hasattr(v, "__len__") else v) for k, v in o.items() if k != "source"})
