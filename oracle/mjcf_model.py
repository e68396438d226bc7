"""ORACLE (test infrastructure, not product code): MJCF -> flat model arrays in numpy.

Independent Python restatement of the subset of MuJoCo's model compiler that the
reference's three scenes need (reference assets/main.xml, ur3e_2f85.xml, ur3e_raw.xml;
SURVEY App. A).  The product's loader is the C++ one in ur3e_b200/csrc/mjcf.cpp; tests
compare its arrays against these.  Array names follow MuJoCo's mjModel.

Parity status: UNPINNED against a real MuJoCo build (mujoco==3.3.3 is not installable
here); pinned only by the tcp@'down' golden vector (reference
assets/main.xml:415) and the nq/nv/nu comments (reference controller/move_l_mug.py:29-32).
"""
import math
import xml.etree.ElementTree as ET

import numpy as np

mjMINVAL = 1e-15

JNT_FREE, JNT_BALL, JNT_SLIDE, JNT_HINGE = 0, 1, 2, 3
GEOM_PLANE, GEOM_HFIELD, GEOM_SPHERE, GEOM_CAPSULE, GEOM_ELLIPSOID, GEOM_CYLINDER, GEOM_BOX, GEOM_MESH = range(8)
GEOM_TYPES = dict(plane=0, hfield=1, sphere=2, capsule=3, ellipsoid=4, cylinder=5, box=6, mesh=7)
EQ_CONNECT, EQ_WELD, EQ_JOINT = 0, 1, 2
TRN_JOINT, TRN_TENDON = 0, 3

DEFAULT_SOLREF = [0.02, 1.0]
DEFAULT_SOLIMP = [0.9, 0.95, 0.001, 0.5, 2.0]

# mesh-only bodies of the ORIGINAL reference XMLs (meshes are absent); same
# numbers as tools/make_assets.py.  Used when a body has no <inertial> and no primitive mass.
MESH_INERTIA = {
    "robotiq_base_mount": dict(pos=[0, 0, 0.002], mass=0.0884, diaginertia=[3.2e-05, 3.2e-05, 6.2e-05]),
    "right_silicone_pad": dict(pos=[0, -0.0056, 0.01875], mass=0.0017, diaginertia=[2.0e-07, 2.68e-07, 6.9e-08]),
    "left_silicone_pad": dict(pos=[0, -0.0056, 0.01875], mass=0.0017, diaginertia=[2.0e-07, 2.68e-07, 6.9e-08]),
}


def _floats(s, n=None, default=None):
    if s is None:
        return None if default is None else list(default)
    v = [float(x) for x in s.split()]
    if n is not None and len(v) < n and default is not None:
        v = v + list(default[len(v):])
    return v


def quat_mul(a, b):
    return np.array([
        a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
        a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
        a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
        a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]])


def quat2mat(q):
    w, x, y, z = q
    return np.array([
        [w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z]])


def quat_norm(q):
    q = np.asarray(q, dtype=np.float64)
    n = np.linalg.norm(q)
    return q / n if n > 0 else np.array([1.0, 0, 0, 0])


class _Defaults:
    """Nested <default class=...> tree (MuJoCo modeling: 'Default settings')."""

    def __init__(self):
        self.classes = {"main": {}}
        self.parent = {"main": None}

    def load(self, elem, cls="main"):
        for e in elem:
            if e.tag == "default":
                name = e.get("class")
                self.classes[name] = {}
                self.parent[name] = cls
                self.load(e, name)
            else:
                self.classes[cls].setdefault(e.tag, {}).update(e.attrib)

    def resolve(self, tag, elem, childclass):
        cls = elem.get("class") or childclass or "main"
        chain = []
        c = cls
        while c is not None:
            chain.append(c)
            c = self.parent[c]
        out = {}
        for c in reversed(chain):
            if tag == "motor":  # actuator shortcuts share the <general> defaults
                out.update(self.classes[c].get("general", {}))
            out.update(self.classes[c].get(tag, {}))
        out.update(elem.attrib)
        return out


def load_mjcf(path):
    """Parse `path` and return a dict of numpy arrays + python lists (names)."""
    root = ET.parse(path).getroot()
    comp = root.find("compiler")
    autolimits = True
    if comp is not None:
        assert comp.get("angle", "degree") == "radian", "only angle=radian models are supported"
        autolimits = comp.get("autolimits", "true") == "true"
    opt = dict(timestep=0.002, gravity=[0, 0, -9.81], impratio=1.0, cone="pyramidal",
               tolerance=1e-8, iterations=100, ls_iterations=50, ls_tolerance=0.01)
    o = root.find("option")
    if o is not None:
        for k in ("timestep", "impratio", "tolerance", "ls_tolerance"):
            if o.get(k) is not None:
                opt[k] = float(o.get(k))
        for k in ("iterations", "ls_iterations"):
            if o.get(k) is not None:
                opt[k] = int(o.get(k))
        if o.get("gravity") is not None:
            opt["gravity"] = _floats(o.get("gravity"))
        if o.get("cone") is not None:
            opt["cone"] = o.get("cone")
    dfl = _Defaults()
    for d in root.findall("default"):
        dfl.load(d, "main")

    B = dict(name=[], parent=[], pos=[], quat=[], ipos=[], iquat=[], mass=[], inertia=[], jntadr=[], jntnum=[])
    J = dict(name=[], type=[], body=[], pos=[], axis=[], range=[], limited=[], armature=[], damping=[],
             frictionloss=[], stiffness=[], springref=[], ref=[], solref=[], solimp=[], margin=[])
    G = dict(name=[], type=[], body=[], pos=[], quat=[], size=[], contype=[], conaffinity=[], condim=[],
             priority=[], friction=[], solref=[], solimp=[], solmix=[], margin=[], gap=[])
    S = dict(name=[], body=[], pos=[], quat=[], size=[])

    def add_body(elem, parent, childclass):
        bid = len(B["name"])
        B["name"].append(elem.get("name", "world" if parent < 0 else "body%d" % bid))
        B["parent"].append(max(parent, 0))
        B["pos"].append(_floats(elem.get("pos"), 3, [0, 0, 0]))
        B["quat"].append(quat_norm(_floats(elem.get("quat"), 4, [1, 0, 0, 0])))
        cc = elem.get("childclass") or childclass
        B["jntadr"].append(len(J["name"]))
        nj = 0
        geom_mass = []  # (mass, pos, quat, diag inertia) of primitive geoms with mass
        for e in elem:
            if e.tag in ("joint", "freejoint"):
                a = dfl.resolve("joint", e, cc) if e.tag == "joint" else dict(e.attrib, type="free")
                jt = dict(free=JNT_FREE, ball=JNT_BALL, slide=JNT_SLIDE, hinge=JNT_HINGE)[a.get("type", "hinge")]
                J["name"].append(a.get("name", ""))
                J["type"].append(jt)
                J["body"].append(bid)
                J["pos"].append(_floats(a.get("pos"), 3, [0, 0, 0]))
                ax = np.array(_floats(a.get("axis"), 3, [0, 0, 1]))
                J["axis"].append(ax / np.linalg.norm(ax))
                rng = _floats(a.get("range"), 2, [0, 0])
                lim = a.get("limited", "auto")
                limited = (lim == "true") or (lim == "auto" and autolimits and a.get("range") is not None)
                J["range"].append(rng)
                J["limited"].append(1 if (limited and jt != JNT_FREE) else 0)
                for k in ("armature", "damping", "frictionloss", "stiffness", "springref", "ref", "margin"):
                    J[k].append(float(a.get(k, 0)))
                J["solref"].append(_floats(a.get("solreflimit"), 2, DEFAULT_SOLREF))
                J["solimp"].append(_floats(a.get("solimplimit"), 5, DEFAULT_SOLIMP))
                nj += 1
            elif e.tag == "geom":
                a = dfl.resolve("geom", e, cc)
                gt = GEOM_TYPES[a.get("type", "sphere")]
                if a.get("mesh") is not None:
                    gt = GEOM_MESH
                G["name"].append(a.get("name", ""))
                G["type"].append(gt)
                G["body"].append(bid)
                gpos = _floats(a.get("pos"), 3, [0, 0, 0])
                gquat = quat_norm(_floats(a.get("quat"), 4, [1, 0, 0, 0]))
                G["pos"].append(gpos)
                G["quat"].append(gquat)
                size = _floats(a.get("size"), 3, [0, 0, 0])
                G["size"].append(size)
                G["contype"].append(int(a.get("contype", 1)))
                G["conaffinity"].append(int(a.get("conaffinity", 1)))
                G["condim"].append(int(a.get("condim", 3)))
                G["priority"].append(int(a.get("priority", 0)))
                G["friction"].append(_floats(a.get("friction"), 3, [1, 0.005, 0.0001]))
                G["solref"].append(_floats(a.get("solref"), 2, DEFAULT_SOLREF))
                G["solimp"].append(_floats(a.get("solimp"), 5, DEFAULT_SOLIMP))
                G["solmix"].append(float(a.get("solmix", 1)))
                G["margin"].append(float(a.get("margin", 0)))
                G["gap"].append(float(a.get("gap", 0)))
                if gt == GEOM_BOX:
                    vol = 8 * size[0] * size[1] * size[2]
                    m = float(a["mass"]) if a.get("mass") is not None else float(a.get("density", 1000)) * vol
                    if m > 0:
                        I = [m / 3 * (size[1] ** 2 + size[2] ** 2), m / 3 * (size[0] ** 2 + size[2] ** 2),
                             m / 3 * (size[0] ** 2 + size[1] ** 2)]
                        geom_mass.append((m, np.array(gpos), gquat, np.array(I)))
            elif e.tag == "site":
                a = dfl.resolve("site", e, cc)
                S["name"].append(a.get("name", ""))
                S["body"].append(bid)
                S["pos"].append(_floats(a.get("pos"), 3, [0, 0, 0]))
                S["quat"].append(quat_norm(_floats(a.get("quat"), 4, [1, 0, 0, 0])))
                S["size"].append(_floats(a.get("size"), 3, [0.005, 0.005, 0.005]))
        B["jntnum"].append(nj)
        inert = elem.find("inertial")
        if inert is None and B["name"][bid] in MESH_INERTIA and not geom_mass and parent >= 0:
            mi = MESH_INERTIA[B["name"][bid]]
            B["mass"].append(mi["mass"]); B["ipos"].append(list(mi["pos"]))
            B["iquat"].append(np.array([1.0, 0, 0, 0])); B["inertia"].append(list(mi["diaginertia"]))
        elif inert is not None:
            B["mass"].append(float(inert.get("mass")))
            B["ipos"].append(_floats(inert.get("pos"), 3, [0, 0, 0]))
            B["iquat"].append(quat_norm(_floats(inert.get("quat"), 4, [1, 0, 0, 0])))
            B["inertia"].append(_floats(inert.get("diaginertia"), 3))
        elif geom_mass:
            # inertia from primitive geoms: parallel-axis composition in the body frame
            mtot = sum(g[0] for g in geom_mass)
            com = sum(g[0] * g[1] for g in geom_mass) / mtot
            I = np.zeros((3, 3))
            for (m, p, q, Id) in geom_mass:
                R = quat2mat(q)
                d = p - com
                I += R @ np.diag(Id) @ R.T + m * (d @ d * np.eye(3) - np.outer(d, d))
            w, V = np.linalg.eigh(I)
            order = np.argsort(-w)  # MuJoCo sorts principal inertias in decreasing order
            w, V = w[order], V[:, order]
            if np.linalg.det(V) < 0:
                V[:, 2] = -V[:, 2]
            if np.allclose(I, np.diag(np.diag(I)), atol=1e-15):
                # already diagonal: keep the body axes (avoids an arbitrary eigenbasis)
                w, V = np.diag(I).copy(), np.eye(3)
            B["mass"].append(mtot); B["ipos"].append(list(com))
            B["iquat"].append(_mat2quat(V)); B["inertia"].append(list(w))
        else:
            B["mass"].append(0.0); B["ipos"].append([0, 0, 0])
            B["iquat"].append(np.array([1.0, 0, 0, 0])); B["inertia"].append([0, 0, 0])
        for e in elem:
            if e.tag == "body":
                add_body(e, bid, cc)
        return bid

    wb = root.find("worldbody")
    add_body(wb, -1, None)

    nbody, njnt = len(B["name"]), len(J["name"])
    m = {"opt": opt, "names": dict(body=B["name"], joint=J["name"], geom=G["name"], site=S["name"])}
    m["nbody"], m["njnt"], m["ngeom"], m["nsite"] = nbody, njnt, len(G["name"]), len(S["name"])
    f64 = lambda x, shape: np.asarray(x, dtype=np.float64).reshape(shape)
    i32 = lambda x: np.asarray(x, dtype=np.int32)
    m["body_parentid"] = i32(B["parent"])
    m["body_pos"] = f64(B["pos"], (nbody, 3)); m["body_quat"] = f64(B["quat"], (nbody, 4))
    m["body_ipos"] = f64(B["ipos"], (nbody, 3)); m["body_iquat"] = f64(B["iquat"], (nbody, 4))
    m["body_mass"] = f64(B["mass"], (nbody,)); m["body_inertia"] = f64(B["inertia"], (nbody, 3))
    m["body_jntadr"] = i32(B["jntadr"]); m["body_jntnum"] = i32(B["jntnum"])
    for b in range(nbody):
        if B["jntnum"][b] == 0:
            m["body_jntadr"][b] = -1

    # joints -> qpos/dof addresses
    qadr, dadr, nq, nv = [], [], 0, 0
    for j in range(njnt):
        qadr.append(nq); dadr.append(nv)
        t = J["type"][j]
        nq += {JNT_FREE: 7, JNT_BALL: 4}.get(t, 1)
        nv += {JNT_FREE: 6, JNT_BALL: 3}.get(t, 1)
    m["nq"], m["nv"] = nq, nv
    m["jnt_type"] = i32(J["type"]); m["jnt_bodyid"] = i32(J["body"])
    m["jnt_qposadr"] = i32(qadr); m["jnt_dofadr"] = i32(dadr)
    m["jnt_pos"] = f64(J["pos"], (njnt, 3)); m["jnt_axis"] = f64(J["axis"], (njnt, 3))
    m["jnt_range"] = f64(J["range"], (njnt, 2)); m["jnt_limited"] = i32(J["limited"])
    m["jnt_stiffness"] = f64(J["stiffness"], (njnt,)); m["jnt_margin"] = f64(J["margin"], (njnt,))
    m["jnt_solref"] = f64(J["solref"], (njnt, 2)); m["jnt_solimp"] = f64(J["solimp"], (njnt, 5))
    qpos0 = np.zeros(nq); qspring = np.zeros(nq)
    dof_body, dof_jnt, dof_parent = [], [], []
    dof_arm, dof_damp, dof_fl = [], [], []
    body_dofadr = -np.ones(nbody, dtype=np.int32); body_dofnum = np.zeros(nbody, dtype=np.int32)
    last_dof_of_body = -np.ones(nbody, dtype=np.int32)
    for j in range(njnt):
        t, b = J["type"][j], J["body"][j]
        if t == JNT_FREE:
            qpos0[qadr[j]:qadr[j] + 3] = B["pos"][b]
            qpos0[qadr[j] + 3:qadr[j] + 7] = B["quat"][b]
            qspring[qadr[j]:qadr[j] + 7] = qpos0[qadr[j]:qadr[j] + 7]
            nd = 6
        else:
            qpos0[qadr[j]] = J["ref"][j]
            qspring[qadr[j]] = J["springref"][j]
            nd = 1
        for k in range(nd):
            d = dadr[j] + k
            if body_dofadr[b] < 0:
                body_dofadr[b] = d
            body_dofnum[b] += 1
            # parent dof: previous dof in the same body, else last dof of nearest ancestor with dofs
            if last_dof_of_body[b] >= 0:
                par = last_dof_of_body[b]
            else:
                par, a = -1, B["parent"][b]
                while True:
                    if last_dof_of_body[a] >= 0:
                        par = last_dof_of_body[a]; break
                    if a == 0:
                        break
                    a = B["parent"][a]
            dof_parent.append(par); last_dof_of_body[b] = d
            dof_body.append(b); dof_jnt.append(j)
            dof_arm.append(J["armature"][j]); dof_damp.append(J["damping"][j]); dof_fl.append(J["frictionloss"][j])
    m["qpos0"], m["qpos_spring"] = qpos0, qspring
    m["dof_bodyid"], m["dof_jntid"], m["dof_parentid"] = i32(dof_body), i32(dof_jnt), i32(dof_parent)
    m["dof_armature"], m["dof_damping"], m["dof_frictionloss"] = f64(dof_arm, (nv,)), f64(dof_damp, (nv,)), f64(dof_fl, (nv,))
    m["dof_solref"] = np.tile(np.array(DEFAULT_SOLREF), (nv, 1)); m["dof_solimp"] = np.tile(np.array(DEFAULT_SOLIMP), (nv, 1))
    m["body_dofadr"], m["body_dofnum"] = body_dofadr, body_dofnum
    # rootid / weldid
    rootid = np.zeros(nbody, dtype=np.int32); weldid = np.zeros(nbody, dtype=np.int32)
    for b in range(1, nbody):
        p = B["parent"][b]
        rootid[b] = b if p == 0 else rootid[p]
        weldid[b] = b if B["jntnum"][b] > 0 else weldid[p]
    m["body_rootid"], m["body_weldid"] = rootid, weldid

    ng = m["ngeom"]
    m["geom_type"] = i32(G["type"]); m["geom_bodyid"] = i32(G["body"])
    m["geom_pos"] = f64(G["pos"], (ng, 3)); m["geom_quat"] = f64(G["quat"], (ng, 4)); m["geom_size"] = f64(G["size"], (ng, 3))
    m["geom_contype"] = i32(G["contype"]); m["geom_conaffinity"] = i32(G["conaffinity"])
    m["geom_condim"] = i32(G["condim"]); m["geom_priority"] = i32(G["priority"])
    m["geom_friction"] = f64(G["friction"], (ng, 3)); m["geom_solref"] = f64(G["solref"], (ng, 2))
    m["geom_solimp"] = f64(G["solimp"], (ng, 5)); m["geom_solmix"] = f64(G["solmix"], (ng,))
    m["geom_margin"] = f64(G["margin"], (ng,)); m["geom_gap"] = f64(G["gap"], (ng,))
    ns = m["nsite"]
    m["site_bodyid"] = i32(S["body"]); m["site_pos"] = f64(S["pos"], (ns, 3)); m["site_quat"] = f64(S["quat"], (ns, 4)); m["site_size"] = f64(S["size"], (ns, 3))

    # tendons (fixed only)
    ten = dict(name=[], adr=[], num=[], jnt=[], coef=[])
    t = root.find("tendon")
    if t is not None:
        for fx in t.findall("fixed"):
            ten["name"].append(fx.get("name", "")); ten["adr"].append(len(ten["jnt"]))
            n = 0
            for jj in fx.findall("joint"):
                ten["jnt"].append(J["name"].index(jj.get("joint"))); ten["coef"].append(float(jj.get("coef"))); n += 1
            ten["num"].append(n)
    m["ntendon"] = len(ten["name"]); m["names"]["tendon"] = ten["name"]
    m["tendon_adr"], m["tendon_num"] = i32(ten["adr"]), i32(ten["num"])
    m["wrap_jnt"], m["wrap_coef"] = i32(ten["jnt"]), f64(ten["coef"], (len(ten["coef"]),))

    # equality
    E = dict(type=[], o1=[], o2=[], data=[], solref=[], solimp=[])
    eq = root.find("equality")
    if eq is not None:
        for e in eq:
            data = np.zeros(11)
            if e.tag == "connect":
                E["type"].append(EQ_CONNECT)
                E["o1"].append(B["name"].index(e.get("body1")))
                E["o2"].append(B["name"].index(e.get("body2")) if e.get("body2") else 0)
                data[0:3] = _floats(e.get("anchor"))
            elif e.tag == "joint":
                E["type"].append(EQ_JOINT)
                E["o1"].append(J["name"].index(e.get("joint1")))
                E["o2"].append(J["name"].index(e.get("joint2")) if e.get("joint2") else -1)
                data[0:5] = _floats(e.get("polycoef"), 5, [0, 1, 0, 0, 0])
            else:
                raise NotImplementedError("equality type %s" % e.tag)
            E["data"].append(data)
            E["solref"].append(_floats(e.get("solref"), 2, DEFAULT_SOLREF))
            E["solimp"].append(_floats(e.get("solimp"), 5, DEFAULT_SOLIMP))
    neq = len(E["type"]); m["neq"] = neq
    m["eq_type"], m["eq_obj1id"], m["eq_obj2id"] = i32(E["type"]), i32(E["o1"]), i32(E["o2"])
    m["eq_data"] = f64(E["data"], (neq, 11)); m["eq_solref"] = f64(E["solref"], (neq, 2)); m["eq_solimp"] = f64(E["solimp"], (neq, 5))

    # actuators
    A = dict(name=[], trntype=[], trnid=[], gain=[], bias=[], ctrlrange=[], ctrllimited=[], forcerange=[],
             forcelimited=[], gear=[])
    act = root.find("actuator")
    if act is not None:
        for e in act:
            a = dfl.resolve(e.tag, e, None)
            A["name"].append(a.get("name", ""))
            if a.get("joint") is not None:
                A["trntype"].append(TRN_JOINT); A["trnid"].append(J["name"].index(a.get("joint")))
            else:
                A["trntype"].append(TRN_TENDON); A["trnid"].append(ten["name"].index(a.get("tendon")))
            if e.tag == "motor":
                A["gain"].append(1.0); A["bias"].append([0, 0, 0])
            elif e.tag == "general":
                A["gain"].append(_floats(a.get("gainprm"), 1, [1])[0])
                bp = _floats(a.get("biasprm"), 3, [0, 0, 0])[:3]
                A["bias"].append(bp if a.get("biastype", "none") == "affine" else [0, 0, 0])
            else:
                raise NotImplementedError("actuator %s" % e.tag)
            cr = a.get("ctrlrange"); fr = a.get("forcerange")
            A["ctrlrange"].append(_floats(cr, 2, [0, 0])); A["forcerange"].append(_floats(fr, 2, [0, 0]))
            cl = a.get("ctrllimited", "auto"); fl = a.get("forcelimited", "auto")
            A["ctrllimited"].append(1 if cl == "true" or (cl == "auto" and autolimits and cr is not None) else 0)
            A["forcelimited"].append(1 if fl == "true" or (fl == "auto" and autolimits and fr is not None) else 0)
            A["gear"].append(_floats(a.get("gear"), 1, [1])[0])
    nu = len(A["name"]); m["nu"] = nu; m["names"]["actuator"] = A["name"]
    m["actuator_trntype"], m["actuator_trnid"] = i32(A["trntype"]), i32(A["trnid"])
    m["actuator_gainprm"] = f64(A["gain"], (nu,)); m["actuator_biasprm"] = f64(A["bias"], (nu, 3))
    m["actuator_ctrlrange"] = f64(A["ctrlrange"], (nu, 2)); m["actuator_ctrllimited"] = i32(A["ctrllimited"])
    m["actuator_forcerange"] = f64(A["forcerange"], (nu, 2)); m["actuator_forcelimited"] = i32(A["forcelimited"])
    m["actuator_gear"] = f64(A["gear"], (nu,))

    # contact excludes / explicit pairs
    excl, pairs = [], []
    c = root.find("contact")
    if c is not None:
        for e in c.findall("exclude"):
            excl.append((B["name"].index(e.get("body1")), B["name"].index(e.get("body2"))))
        for e in c.findall("pair"):
            g1, g2 = e.get("geom1"), e.get("geom2")
            if g1 in G["name"] and g2 in G["name"]:
                pairs.append((G["name"].index(g1), G["name"].index(g2), e.attrib))
            # pairs naming a geom absent from the model (mesh-stripped) are skipped
    m["exclude"] = excl
    m["pair_candidates"] = _collision_pairs(m, excl, pairs)

    # torque sensors: the sites they are attached to, in sensor order
    sn = root.find("sensor")
    m["sensor_torque_site"] = i32([S["name"].index(e.get("site")) for e in sn.findall("torque")] if sn is not None else [])

    # keyframes
    K = dict(name=[], qpos=[], qvel=[])
    kf = root.find("keyframe")
    if kf is not None:
        for e in kf.findall("key"):
            K["name"].append(e.get("name", ""))
            K["qpos"].append(_floats(e.get("qpos"), nq, qpos0) if e.get("qpos") else list(qpos0))
            K["qvel"].append(_floats(e.get("qvel"), nv, np.zeros(nv)) if e.get("qvel") else [0.0] * nv)
    m["nkey"] = len(K["name"]); m["names"]["key"] = K["name"]
    m["key_qpos"] = f64(K["qpos"], (m["nkey"], nq)); m["key_qvel"] = f64(K["qvel"], (m["nkey"], nv))
    return m


def _mat2quat(R):
    tr = R[0, 0] + R[1, 1] + R[2, 2]
    if tr > 0:
        s = math.sqrt(tr + 1.0) * 2
        q = [0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s]
    elif R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
        s = math.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2]) * 2
        q = [(R[2, 1] - R[1, 2]) / s, 0.25 * s, (R[0, 1] + R[1, 0]) / s, (R[0, 2] + R[2, 0]) / s]
    elif R[1, 1] > R[2, 2]:
        s = math.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2]) * 2
        q = [(R[0, 2] - R[2, 0]) / s, (R[0, 1] + R[1, 0]) / s, 0.25 * s, (R[1, 2] + R[2, 1]) / s]
    else:
        s = math.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1]) * 2
        q = [(R[1, 0] - R[0, 1]) / s, (R[0, 2] + R[2, 0]) / s, (R[1, 2] + R[2, 1]) / s, 0.25 * s]
    return quat_norm(q)


def _collision_pairs(m, excl, pairs):
    """Static candidate list of primitive geom pairs (SURVEY App. B.9 filter rules):
    contype/conaffinity, not same weld body, not parent-child (unless the parent is
    static), not <exclude>; explicit <pair>s always kept and override the mixing.
    Each row: g1, g2, condim, friction[3 used: slide, slide, (torsion...)], solref[2],
    solimp[5], margin, gap.  Mesh geoms never collide here; plane-plane and
    same-body pairs are dropped."""
    ng = m["ngeom"]
    out = []
    explicit = {}
    for (g1, g2, attr) in pairs:
        explicit[(min(g1, g2), max(g1, g2))] = (g1, g2, attr)
    exset = {(min(a, b), max(a, b)) for a, b in excl}
    parent, weld = m["body_parentid"], m["body_weldid"]
    for g1 in range(ng):
        for g2 in range(g1 + 1, ng):
            t1, t2 = m["geom_type"][g1], m["geom_type"][g2]
            if t1 == GEOM_MESH or t2 == GEOM_MESH:
                continue
            if t1 == GEOM_PLANE and t2 == GEOM_PLANE:
                continue
            if t1 not in (GEOM_PLANE, GEOM_BOX) or t2 not in (GEOM_PLANE, GEOM_BOX):
                raise NotImplementedError("only plane/box primitives are supported")
            b1, b2 = m["geom_bodyid"][g1], m["geom_bodyid"][g2]
            key = (g1, g2)
            if key not in explicit:
                w1, w2 = weld[b1], weld[b2]
                if w1 == w2:
                    continue  # same body or welded together (both static included)
                if not ((m["geom_contype"][g1] & m["geom_conaffinity"][g2]) or (m["geom_contype"][g2] & m["geom_conaffinity"][g1])):
                    continue
                # parent-child filter on weld bodies (skipped when the parent is the static world weld 0)
                pw1 = weld[parent[w1]] if w1 != 0 else -1
                pw2 = weld[parent[w2]] if w2 != 0 else -1
                if (w1 != 0 and w2 != 0) and (pw1 == w2 or pw2 == w1):
                    continue
                if (min(b1, b2), max(b1, b2)) in exset:
                    continue
            p1, p2 = m["geom_priority"][g1], m["geom_priority"][g2]
            if p1 != p2:
                gw = g1 if p1 > p2 else g2
                fr = m["geom_friction"][gw]
                solref, solimp, condim = m["geom_solref"][gw], m["geom_solimp"][gw], int(m["geom_condim"][gw])
            else:
                fr = np.maximum(m["geom_friction"][g1], m["geom_friction"][g2])
                s1, s2 = m["geom_solmix"][g1], m["geom_solmix"][g2]
                mix = s1 / (s1 + s2) if (s1 + s2) > mjMINVAL else 0.5
                solref = mix * m["geom_solref"][g1] + (1 - mix) * m["geom_solref"][g2]
                solimp = mix * m["geom_solimp"][g1] + (1 - mix) * m["geom_solimp"][g2]
                condim = int(max(m["geom_condim"][g1], m["geom_condim"][g2]))
            row = dict(g1=g1, g2=g2, condim=condim,
                       friction=[fr[0], fr[0], fr[1], fr[2], fr[2]],
                       solref=list(solref), solimp=list(solimp),
                       margin=float(max(m["geom_margin"][g1], m["geom_margin"][g2])),
                       gap=float(max(m["geom_gap"][g1], m["geom_gap"][g2])))
            if key in explicit:
                # explicit <pair>: never filtered; attributes it does not set are inferred from its geoms (mjCPair::Compile)
                e1, e2, attr = explicit[key]
                row["g1"], row["g2"] = e1, e2
                if "condim" in attr: row["condim"] = int(attr["condim"])
                if "friction" in attr: row["friction"] = _floats(attr.get("friction"), 5, row["friction"])
                if "solref" in attr: row["solref"] = _floats(attr.get("solref"), 2, row["solref"])
                if "solimp" in attr: row["solimp"] = _floats(attr.get("solimp"), 5, row["solimp"])
                if "margin" in attr: row["margin"] = float(attr["margin"])
                if "gap" in attr: row["gap"] = float(attr["gap"])
            out.append(row)
    # canonical order: plane first within a pair, list sorted by (g1, g2)
    for r in out:
        if m["geom_type"][r["g2"]] == GEOM_PLANE:
            r["g1"], r["g2"] = r["g2"], r["g1"]
    out.sort(key=lambda r: (min(r["g1"], r["g2"]), max(r["g1"], r["g2"])))
    return out
