"""JSON report, unchanged format.

Host-side mirror of /root/reference/src/report.go:46-148 (title -> {Xl, Yl, Vis, Pos, Data{str -> num}}),
fragstats.go:120-141 (the nine FragStats panels) and thermocycler.go:195-221 (efficiency panels).
tools/plot_rlsim_report:144-155 is the consumer of this schema.
"""
import json


def _g(x):  # Go's %g for the float keys used by the reference (0, 0.5, 1, ...)
    s = "%g" % x
    return s


class Report:
    def __init__(self, path):
        self.path = path
        self.pool = {}
        self.cursor = 0

    def _add(self, data, xl, yl, title, vis):
        self.pool[title] = {"Xl": xl, "Yl": yl, "Vis": vis, "Pos": self.cursor, "Data": data}
        self.cursor += 1

    def report_map(self, m, xl, yl, title, vis):  # ReportMapInt32t64 / ReportSliceInt32t64 / ReportSliceInt32f64
        self._add({str(int(k)): v for k, v in m.items()}, xl, yl, title, vis)

    def report_float(self, xs, ys, xl, yl, title, vis):  # ReportSliceFloat64f64
        self._add({_g(x): y for x, y in zip(xs, ys)}, xl, yl, title, vis)

    def write_json(self):
        with open(self.path, "w") as f:
            json.dump(self.pool, f, separators=(",", ":"), sort_keys=True)  # Go marshals map keys sorted
