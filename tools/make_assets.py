#!/usr/bin/env python3
"""Generate the mesh-stripped MJCF models this repo simulates.

The reference ships its scenes as MJCF (assets/main.xml, ur3e_2f85.xml, ur3e_raw.xml)
but git-ignores the 15 STL + 1 OBJ meshes they name (reference .gitignore:2-4), so the
files cannot be compiled as shipped.  This script derives, from the
reference XML, the model that CAN be compiled anywhere (by our loader and by a stock
MuJoCo alike):

  * <asset>, <visual>, lights and every mesh geom are dropped (no collision, no mass);
  * bodies whose mass came only from mesh geoms get an explicit <inertial> taken from
    MESH_INERTIA below (documented estimates, DESIGN.md "mesh inertia");
  * <pair>s that name a removed geom are dropped;
  * comments / dead alternatives are dropped, numbers are kept verbatim.

Run in the build container only (needs /root/reference):
    python tools/make_assets.py            # writes ur3e_b200/assets/*.xml
The outputs are committed; nothing at run time reads /root/reference.
"""
import os
import sys
import xml.etree.ElementTree as ET

REF = os.environ.get("UR3E_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "ur3e_b200", "assets")

# Bodies that have no <inertial> and only mesh geoms (reference assets/main.xml:152-155,
# 197-199, 235-237).  MuJoCo would integrate the (absent) STL volumes at density 1000.
# Estimates: base_mount = 75 mm dia x 10 mm coupling disc, counted twice (visual +
# collision geom both contribute mass); silicone pad = 22 x 2 x 37.5 mm sheet.
MESH_INERTIA = {
    "robotiq_base_mount": dict(pos="0 0 0.002", mass="0.0884",
                               diaginertia="3.2e-05 3.2e-05 6.2e-05"),
    "right_silicone_pad": dict(pos="0 -0.0056 0.01875", mass="0.0017",
                               diaginertia="2.0e-07 2.68e-07 6.9e-08"),
    "left_silicone_pad": dict(pos="0 -0.0056 0.01875", mass="0.0017",
                              diaginertia="2.0e-07 2.68e-07 6.9e-08"),
}

# Box hulls standing in for the collision meshes of the arm links and the gripper base (main.xml only: the scene of the gym envs).
# The reference collides those meshes with the table, the mug and each other (dynamic pairs + its explicit <pair>s,
# assets/main.xml:325-334) and reads the resulting contacts in get_self_collision / get_table_collision
# (utils/gym_utils.py:146-201).  The STL files are not in the repository, so the dimensions below are estimates from
# the link offsets in the XML and the UR3e / 2F-85 data sheets (tube radii 40 / 33 / 32 mm, base 55 mm), kept slightly inside the
# real hulls so that no pose near keyframe 'down' reports a false self-collision.  geom name -> (body, pos, half-sizes).
# Collision classes: proxies use contype = conaffinity = 2, the table plane and the mug 3, the pad boxes keep the reference's 1:
# proxies collide with the table, the mug and each other (non-adjacent links) but not with the pad boxes, which bounds the candidate
# pair list (48 pairs) -- pad-vs-arm-link contacts are the one family of the reference's dynamic pairs that is not modelled.
PROXIES = {
    "base":      ("robot_base",     "0 0 0.045",   "0.055 0.055 0.045"),
    "shoulder":  ("shoulder_link",  "0 0.015 0",   "0.045 0.06 0.06"),
    "upperarm":  ("upper_arm_link", "0 0 0.122",   "0.04 0.04 0.162"),
    "forearm":   ("forearm_link",   "0 0 0.1065",  "0.033 0.033 0.1395"),
    "wrist1":    ("wrist_1_link",   "0 0.015 0",   "0.032 0.055 0.032"),
    "wrist2":    ("wrist_2_link",   "0 0 0.02",    "0.032 0.032 0.05"),
    "wrist3":    ("wrist_3_link",   "0 0.03 0",    "0.032 0.05 0.032"),
    "collision": ("gripper_base",   "0 0 0.045",   "0.017 0.036 0.045"),
}
PROXY_BITS, SHARED_BITS = "2", "3"

MESH_CLASSES = {"visual", "collision"}  # default classes whose geoms are type="mesh"


def is_mesh_geom(g):
    if g.get("type") == "mesh" or g.get("mesh") is not None:
        return True
    return g.get("class") in MESH_CLASSES and g.get("type") is None


def add_proxies(root):
    """main.xml: box hulls for the stripped collision meshes (see PROXIES); must run before the <pair> pruning."""
    bodies = {b.get("name"): b for b in root.iter("body")}
    for name, (body, pos, size) in PROXIES.items():
        g = ET.Element("geom", dict(name=name, type="box", pos=pos, size=size, mass="0", contype=PROXY_BITS, conaffinity=PROXY_BITS))
        b = bodies[body]
        idx = max([i for i, e in enumerate(list(b)) if e.tag in ("inertial", "joint", "site")] + [-1]) + 1
        b.insert(idx, g)
    for g in root.iter("geom"):
        if g.get("name") in ("table", "fish"):
            g.set("contype", SHARED_BITS); g.set("conaffinity", SHARED_BITS)


def strip(root, proxies=False):
    removed = set()
    for tag in ("asset", "visual", "size"):
        for e in root.findall(tag):
            root.remove(e)
    # defaults: drop <mesh .../> defaults and cosmetic attributes
    for d in root.iter("default"):
        for e in list(d):
            if e.tag == "mesh":
                d.remove(e)
    for parent in root.iter():
        for e in list(parent):
            if e.tag == "light":
                parent.remove(e)
            elif e.tag == "geom" and is_mesh_geom(e):
                if e.get("name"):
                    removed.add(e.get("name"))
                parent.remove(e)
    for e in root.iter():
        for a in ("rgba", "material", "group"):
            if a in e.attrib:
                del e.attrib[a]
    if proxies:
        add_proxies(root)
        removed -= set(PROXIES)
    for body in root.iter("body"):
        name = body.get("name")
        if name in MESH_INERTIA and body.find("inertial") is None:
            body.insert(0, ET.Element("inertial", MESH_INERTIA[name]))
    for c in root.findall("contact"):
        for p in list(c.findall("pair")):
            if p.get("geom1") in removed or p.get("geom2") in removed:
                c.remove(p)
    comp = root.find("compiler")
    if comp is not None:
        for a in ("meshdir", "texturedir"):
            comp.attrib.pop(a, None)
    return removed


def main():
    os.makedirs(OUT, exist_ok=True)
    for name in ("main.xml", "ur3e_2f85.xml", "ur3e_raw.xml"):
        tree = ET.parse(os.path.join(REF, "assets", name))  # ET drops comments
        root = tree.getroot()
        removed = strip(root, proxies=(name == "main.xml"))
        ET.indent(tree, space="  ")
        path = os.path.join(OUT, name)
        with open(path, "w") as f:
            f.write("<!-- generated by tools/make_assets.py from the reference's assets/%s "
                    "(mesh-stripped; see DESIGN.md) -->\n" % name)
            tree.write(f, encoding="unicode")
            f.write("\n")
        print("wrote", path, "removed named mesh geoms:", sorted(removed))


if __name__ == "__main__":
    sys.exit(main())
