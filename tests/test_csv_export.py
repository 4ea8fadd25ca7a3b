"""exportDataToFile (RigidBodySystem.java:495-542): the CSV writer reproduces the reference's own recorded log byte for byte
(layout, ", " separators, "\\n " line ends, Java's Double.toString), checked on a sample of the authors' CSV."""
import os
import types

from adaptivemerging_b200.csvlog import HEADER, CsvLog, format_row, java_double

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_csv_sample.csv")


def test_rows_of_the_reference_log_are_reproduced_byte_for_byte():
    txt = open(GOLDEN, newline="").read()
    parts = txt.split("\n ")
    assert parts[0] == ", ".join(HEADER)
    rows = [p for p in parts[1:] if p]
    assert len(rows) == 40
    for r in rows:
        vals = [float(t) for t in r.split(", ")]
        assert format_row(vals) == r + "\n "


def test_java_double_to_string():
    for x, s in [(0.0, "0.0"), (1.0, "1.0"), (0.001, "0.001"), (9.99e-4, "9.99E-4"), (1e7, "1.0E7"), (9999999.0, "9999999.0"),
                 (4.38629e-4, "4.38629E-4"), (0.0012078540000000002, "0.0012078540000000002"), (123.456, "123.456"), (-2.5e-9, "-2.5E-9"),
                 (1e-3 * 1.5, "0.0015"), (12345678.9, "1.23456789E7"), (100.0, "100.0")]:
        assert java_double(x) == s, (x, java_double(x), s)


def test_stream_protocol(tmp_path):
    """the call that opens the file writes only the header; switching saveCSV off closes it"""
    t = types.SimpleNamespace(n_bodies=328, n_contacts=12, detection=1.5e-3, warmstart=2e-6, lcp_solve=0.0, update_collections=0.0,
                              contact_ordering=0.0, single_it_pgs=0.0, merging=1e-6, merging_build=0.0, unmerging=0.0, unmerging_build=0.0,
                              compute_time=2e-3)
    log = CsvLog()
    name = str(tmp_path / "tower")
    log.export(True, name, True, t)
    log.export(True, name, True, t)
    log.export(False, name, True, t)
    assert log.stream is None
    txt = open(name + "_merged.csv", newline="").read()
    assert txt == ", ".join(HEADER) + "\n " + "328, 12, 0.0015, 2.0E-6, 0.0, 0.0, 0.0, 0.0, 1.0E-6, 0.0, 0.0, 0.0, 0.002\n "
