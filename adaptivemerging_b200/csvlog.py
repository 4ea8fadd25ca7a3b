"""RigidBodySystem.exportDataToFile (RigidBodySystem.java:495-542): the 13-column CSV step log, byte for byte.

Layout quirks of the reference that are kept: fields are separated by ", ", lines end with "\\n " (a newline followed
by a space, so every data row starts with a space), the call that opens the file writes ONLY the header (the step that
switched saveCSV on leaves no row), doubles are printed by Java's Double.toString.
"""
import math

HEADER = ["#bodies", "#contacts", "detection", "warmstart", "LCPSolve", "updateCollections", "contactOrdering", "singleItPGS",
          "merging", "mergingBuild", "unmerging", "unmergingBuild", "computeTime"]


def java_double(x):
    """java.lang.Double.toString: shortest digits that round-trip; decimal notation for 1e-3 <= |x| < 1e7, otherwise
    computerised scientific notation d.dddE[-]n; always at least one digit after the point."""
    x = float(x)
    if math.isnan(x):
        return "NaN"
    if math.isinf(x):
        return "Infinity" if x > 0 else "-Infinity"
    if x == 0.0:
        return "-0.0" if math.copysign(1.0, x) < 0 else "0.0"
    sign = "-" if x < 0 else ""
    mant, exp = repr(abs(x)).lower().partition("e")[::2]  # shortest round-trip digits
    e10 = int(exp) if exp else 0
    ip, _, fp = mant.partition(".")
    digits = (ip + fp).lstrip("0")
    # decimal exponent of the first significant digit
    lead = len(ip.lstrip("0")) - 1 + e10 if ip.strip("0") else -(len(fp) - len(fp.lstrip("0"))) - 1 + e10
    digits = digits.rstrip("0") or "0"
    if 1e-3 <= abs(x) < 1e7:
        if lead >= 0:
            whole, frac = digits[:lead + 1].ljust(lead + 1, "0"), digits[lead + 1:]
        else:
            whole, frac = "0", "0" * (-lead - 1) + digits
        return f"{sign}{whole}.{frac or '0'}"
    return f"{sign}{digits[0]}.{digits[1:] or '0'}E{lead}"


def row_from_timings(t):
    """the 13 fields from an am3d_timings record"""
    return [t.n_bodies, t.n_contacts, t.detection, t.warmstart, t.lcp_solve, t.update_collections, t.contact_ordering,
            t.single_it_pgs, t.merging, t.merging_build, t.unmerging, t.unmerging_build, t.compute_time]


def format_row(values):
    return ", ".join(str(int(v)) if k < 2 else java_double(v) for k, v in enumerate(values)) + "\n "


class CsvLog:
    """stream state of exportDataToFile: None until saveCSV is switched on, header on the opening call, rows afterwards"""

    def __init__(self):
        self.stream = None

    def export(self, save_csv, scene_name, merging_enabled, timings):
        if save_csv:
            if self.stream is None:
                self.stream = open(f"{scene_name}_merged.csv" if merging_enabled else f"{scene_name}.csv", "w", newline="")
                self.stream.write(", ".join(HEADER) + "\n ")
            else:
                self.stream.write(format_row(row_from_timings(timings)))
        elif self.stream is not None:
            self.stream.close()
            self.stream = None

    def close(self):
        if self.stream is not None:
            self.stream.close()
            self.stream = None
