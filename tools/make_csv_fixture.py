"""tests/golden/ref_csv_sample.csv = the first 40 lines of one of the step logs the authors recorded with the Java simulator
(RigidBodySystem.exportDataToFile): reference OUTPUT used to pin the CSV writer's layout and number formatting.
Run in the build container (needs /root/reference)."""
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = "/root/reference/scenes3D/csv/conditional_acceptance_revisions/tower25platform10_merged.csv"
txt = open(src, newline="").read()
parts = txt.split("\n ")
# header + rows 100..139 (contacts appear around row 60, merges around row 110)
sample = "\n ".join([parts[0]] + parts[100:140]) + "\n "
open(os.path.join(ROOT, "tests", "golden", "ref_csv_sample.csv"), "w", newline="").write(sample)
print(len(sample), "bytes")
