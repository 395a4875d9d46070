"""Writes the robot descriptions used by tests and bench.py into tests/data/*.xml.

The parameter values are the ones of the reference's ctr_resources/*.xml (listed in SURVEY.md
§9.2: section A (VC) K=(5,5) kappa=52.2 L=0.06647 r=(0.00125,0.001), A' (CC) K=5, B (CC) K=1
kappa=52.63 L=0.0667 r=0.001, B' (VC) K=(1,1), C (CC) K=0.05 kappa=1e-6 L=0.01265 r=0.0005; nu=0.3,
phi=0.001 everywhere; one shared entry frame).  The files are generated, not copied: V1 schema
(ctr_static.h:132-208) for the four shipped robots and one V2-schema file (ctr_static.h:461-532)
to exercise the second parser.  `python tests/golden/make_robot_xml.py` regenerates them.
"""
import os

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "data")

ENTRY = [[-0.69710, -0.59750, -0.39620, 0.079],
         [-0.69710, 0.69400, 0.18010, 0.077],
         [0.16730, 0.40170, -0.90040, 0.148]]


def tube(k, kappa, r, L, nu=0.3, phi=0.001):
    return dict(stiffness=k, curvature=kappa, radius=r, poissonRatio=nu, phi=phi, length=L)


A_VC = [tube(5, 52.2, 0.00125, 0.06647), tube(5, 52.2, 0.001, 0.06647)]
A_CC = [tube(5, 52.2, 0.00125, 0.06647)]
B_CC = [tube(1, 52.63, 0.001, 0.0667)]
B_VC = [tube(1, 52.63, 0.001, 0.0667), tube(1, 52.63, 0.001, 0.0667)]
C_CC = [tube(0.05, 0.000001, 0.0005, 0.01265)]

ROBOTS = {
    "ctr_sec2_tub3.xml": [A_VC, B_CC],
    "ctr_sec3_tub3.xml": [A_CC, B_CC, C_CC],
    "ctr_sec3_tub4.xml": [A_VC, B_CC, C_CC],
    "ctr_sec3_tub5.xml": [A_VC, B_VC, C_CC],
}


def fmt(v):
    return repr(float(v)) if not float(v).is_integer() else str(int(v))


def write_v1(path, sections):
    with open(path, "w") as f:
        f.write('<?xml version="1.0"?>\n<robot>\n')
        for i, sec in enumerate(sections):
            f.write("  <section>\n    <name>Manipulation %d</name>\n" % (i + 1))
            for t in sec:
                f.write("    <tube>\n")
                for key in ("stiffness", "curvature", "radius", "poissonRatio", "phi", "length"):
                    f.write("      <%s>%s</%s>\n" % (key, fmt(t[key]), key))
                f.write("    </tube>\n")
            f.write("  </section>\n")
        f.write("  <entryFrame>\n    <!-- row-major 3x4 entry frame; not exactly orthogonal on purpose -->\n")
        for r in range(3):
            f.write("    " + " ".join("<x%d%d>%s</x%d%d>" % (r, c, fmt(ENTRY[r][c]), r, c) for c in range(4)) + "\n")
        f.write("    <x30>0</x30> <x31>0</x31> <x32>0</x32> <x33>1</x33>\n  </entryFrame>\n</robot>\n")


def write_v2(path, sections):
    with open(path, "w") as f:
        f.write('<?xml version="1.0"?>\n<robot version="2.0">\n')
        for i, sec in enumerate(sections):
            f.write("  <Section>\n    <name>Manipulation %d</name>\n" % (i + 1))
            f.write("    <length>%s</length>\n    <phi>%s</phi>\n" % (fmt(sec[0]["length"]), fmt(sec[0]["phi"])))
            for t in sec:
                f.write("    <tube>\n")
                for key in ("stiffness", "curvature", "radius", "poissonRatio"):
                    f.write("      <%s>%s</%s>\n" % (key, fmt(t[key]), key))
                f.write("    </tube>\n")
            f.write("  </Section>\n")
        f.write("  <entryFrame>\n")
        for r in range(3):
            f.write("    " + " ".join("<E%d%d>%s</E%d%d>" % (r + 1, c + 1, fmt(ENTRY[r][c]), r + 1, c + 1) for c in range(4)) + "\n")
        f.write("  </entryFrame>\n</robot>\n")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    for name, secs in ROBOTS.items():
        write_v1(os.path.join(OUT, name), secs)
    write_v2(os.path.join(OUT, "ctr_sec3_tub4_v2.xml"), ROBOTS["ctr_sec3_tub4.xml"])
    print("wrote", sorted(os.listdir(OUT)))
