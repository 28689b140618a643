"""Independent pure-Python model of the reference's grouped streaming window (small inputs only).

Written separately from oracle/dnz_oracle.c (dict-of-dicts, row at a time) so the two can check each other.
Follows grouped_window_agg_stream.rs:326-349 (per-batch order), streaming_window.rs:1053-1094 (windows),
utils/time.rs:31-57 (watermark) and the DataFusion-42 accumulator rules listed in SURVEY.md §8a.
"""
import struct

F64_MAX = 1.7976931348623157e308


def total_key(d: float) -> int:
    b = struct.unpack("<q", struct.pack("<d", d))[0]
    return b ^ (((b >> 63) & 0xFFFFFFFFFFFFFFFF) >> 1)


def windows_for(mn, mx, L, S):
    Ls = L // 1000

    def snap(t):
        return (t // 1000 // Ls) * Ls * 1000
    out = []
    if S:
        cur = snap(mn - L)
        while cur <= mx:
            end = cur + L
            if not (mn > end or mx < cur):
                out.append((cur, end))
            cur += S
    else:
        cur = snap(mn)
        while cur <= mx:
            out.append((cur, cur + L))
            cur += L
    return out


class PyModel:
    def __init__(self, L, S=0, filt=None):
        self.L, self.S, self.filt = L, S, filt
        self.frames = {}      # start -> {"end":, "groups": {key: state}, order: [keys]}
        self.wm = None
        self.seq = 0
        self.out = []

    def push(self, rows):
        """rows: list of (ts|None, val|None, key(bytes)|None)"""
        seq = self.seq
        self.seq += 1
        if not rows:
            return
        tss = [r[0] for r in rows if r[0] is not None]
        mn, mx = min(tss), max(tss)
        for (s, e) in windows_for(mn, mx, self.L, self.S):
            fr = self.frames.setdefault(s, {"end": e, "groups": {}})
            for ts, val, key in rows:
                if ts is None or not (s <= ts < e):
                    continue
                st = fr["groups"].setdefault(key, {"cnt": 0, "mn": F64_MAX, "mx": -F64_MAX, "sum": 0.0, "seen": False})
                if val is not None:
                    st["cnt"] += 1
                    st["seen"] = True
                    if st["mn"] > val:
                        st["mn"] = val
                    if st["mx"] < val:
                        st["mx"] = val
                    st["sum"] += val
        if self.wm is None or self.wm <= mn:
            self.wm = mn
        for s in sorted(self.frames):
            fr = self.frames[s]
            if self.wm >= fr["end"]:
                for key, st in fr["groups"].items():
                    if st["seen"]:
                        row = (s, fr["end"], key, st["cnt"], st["mn"], st["mx"], st["sum"] / st["cnt"], seq)
                    else:
                        row = (s, fr["end"], key, st["cnt"], None, None, None, seq)
                    if self._pass(row):
                        self.out.append(row)
                del self.frames[s]

    def _pass(self, row):
        if not self.filt:
            return True
        col, op, lit = self.filt
        v = {"count": row[3], "min": row[4], "max": row[5], "average": row[6]}[col]
        if v is None:
            return False
        a, b = total_key(float(v)), total_key(float(lit))
        return {">": a > b, ">=": a >= b, "<": a < b, "<=": a <= b, "==": a == b, "!=": a != b}[op]
