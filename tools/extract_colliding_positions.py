"""Extracts the 45 cylinder positions of the reference's validation/collidingParticles case
(constant/IBMProperties) into ibmfoam_b200/data/colliding_positions.txt.  Run in the build container
only (needs /root/reference); the output is committed so nothing reads the reference at run time."""
import re, sys
src = "/root/reference/validation/collidingParticles/constant/IBMProperties"
txt = open(src).read()
pos = re.findall(r"position\s*\(([^)]*)\)", txt)
with open("ibmfoam_b200/data/colliding_positions.txt", "w") as f:
    f.write("# x y z of particle.0 .. particle.44 (validation/collidingParticles/constant/IBMProperties)\n")
    for p in pos:
        f.write(" ".join(p.split()) + "\n")
print(len(pos), "positions")
