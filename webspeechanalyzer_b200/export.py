"""Feature-row export in the web app's own storage format, so that rows extracted on the GPU
load straight into WebSpeechAnalyzer (`Load_JSON_Data`, /root/reference/src/localstore.js:1125) for training.

    call_backed   /root/reference/src/index.js:35-99      which callback becomes which stored rows (per output level)
    StoreFeatures /root/reference/src/localstore.js:39-65  key "<db>#<file>#<si>", expected row width per level (:7)
    JSON export   /root/reference/src/localstore.js:878-887 [{file, seg, time, features, origin, true, pred}, ...]
    CSV export    /root/reference/src/localstore.js:900-990 "file,seg,t0,td,x0,...,xN,\\r\\n" + one line per row

Numbers are printed the way JavaScript prints them (String(x) / JSON.stringify: shortest round-trip digits, integers
without ".0", NaN -> null in JSON and "NaN" in CSV).  Host-side formatting only -- no analysis happens here.
"""
from __future__ import annotations

import json
from decimal import Decimal

import numpy as np

# process_exp_features_len, /root/reference/src/localstore.js:7 (index = output level)
EXPECTED_LEN = [0, 1, 2, 3, 4, 53, 0, 0, 0, 0, 9, 264, 23, 53, 14, 15]


def js_number(x: float) -> str:
    """Number::toString(x) of ECMAScript for a double."""
    x = float(x)
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "Infinity" if x > 0 else "-Infinity"
    if x == int(x) and abs(x) < 1e21:
        return str(int(x))
    r = repr(x)
    if "e" not in r:
        return r
    m, e = r.split("e")
    ex = int(e)
    if m.endswith(".0"):
        m = m[:-2]
    if -7 < ex < 21:
        return format(Decimal(r), "f")
    return m + "e" + ("+" if ex > 0 else "-") + str(abs(ex))


def _json_value(v):
    if isinstance(v, str):
        return json.dumps(v)
    if v is None:
        return "null"
    if isinstance(v, (list, tuple, np.ndarray)):
        return "[" + ",".join(_json_value(x) for x in v) + "]"
    x = float(v)
    return "null" if x != x or x in (float("inf"), float("-inf")) else js_number(x)     # JSON.stringify(NaN) === "null"


def stored_rows(level: int, db_id, file_name: str, calls) -> list[dict]:
    """What call_backed (src/index.js:35) hands to StoreFeatures for the callbacks of one file, in order.

    `calls` = the callback argument tuples (api.segment_callbacks / what LaunchAudioNodes passes to `callback`).
    Returns records {key, file, seg, time, features}; `seg` is String(si) -- `si + ph/100` for syllable rows."""
    out = []

    def store(si, time, row):
        if len(row) != EXPECTED_LEN[level]:            # StoreFeatures refuses rows of the wrong width (localstore.js:46)
            return
        seg = js_number(si)
        out.append({"key": f"{db_id}#{file_name}#{seg}", "file": file_name, "seg": seg, "time": list(time),
                    "features": [float(x) for x in row]})

    for si, _labels, times, payload in calls:
        if level == 11:
            store(0, times, payload)                   # always segment 0: each callback overwrites the clip's row
        elif level in (13, 12):
            for ph, row in enumerate(payload):
                store(si + ph / 100, times[ph], row)
        elif level == 10:
            for syl, frames in enumerate(payload):     # mean of the syllable's Float32Array(9) rows, accumulated in float32
                acc = np.array(frames[0], np.float32).copy()
                for fr in frames[1:]:
                    acc = (acc + np.asarray(fr, np.float32)).astype(np.float32)
                n = np.float32(len(frames))
                acc = np.where(acc != 0, (acc / n).astype(np.float32), acc)
                store(si + syl / 100, times[syl], acc)
        elif level == 5:
            store(si, times, payload)
        # level 4 rows are not stored by the app (src/index.js:94-98)
    if level == 11 and out:
        out = out[-1:]                                 # same localStorage key: the last cumulative row wins
    return out


def to_json(rows: list[dict], origin=None) -> str:
    """The app's "data_<db>.json" (localstore.js:878-887): loadable with Load_JSON_Data (:1125)."""
    items = []
    for r in rows:
        items.append("{" + ",".join([
            '"file":' + json.dumps(r["file"]), '"seg":' + json.dumps(r["seg"]), '"time":' + _json_value(r["time"]),
            '"features":' + _json_value(r["features"]), '"origin":' + _json_value(origin), '"true":null', '"pred":null']) + "}")
    return "[" + ",".join(items) + "]"


def to_csv(rows: list[dict]) -> str:
    """The app's CSV (localstore.js:900-990, horizontal_spread_features, no label columns)."""
    if not rows:
        return "file,seg,t0,td,\r\n"
    n = len(rows[0]["features"])
    lines = ["file,seg,t0,td," + "".join(f"x{i}," for i in range(n)) + "\r\n"]
    for r in rows:
        t0, td = (t if isinstance(t, str) else js_number(t) for t in r["time"])
        lines.append(f"{r['file']},{r['seg']},{t0},{td}," + "".join(js_number(x) + "," for x in r["features"]) + "\r\n")
    return "".join(lines)
