"""TEST INFRASTRUCTURE ONLY.  Executes the REFERENCE'S OWN code -- the minified formantanalyzer@1.1.6 modules inside
/root/reference/dist/main.js (line 2: webpack module 584, inner modules 0 stats @B1065, 3 segmentor @B23403,
4 formants @B31782, 7 utterance @B107866) -- on uint32 frames, with oracle/minijs (this image has no JS engine).

The module sources are read from the reference tree at run time and never copied into this repo; what is
committed are the OUTPUTS (tests/golden/ref_js_*.json, made by tests/golden/make_ref_js_golden.py), which pin the
C oracle on the GPU box where /root/reference does not exist.

Event model: the worklet posts one Uint32Array per MessagePort message, so every
spectrum_push is its own macro-task -- the micro-task queue (the `.then(L(-1), P())` after a finalisation) is drained
after each push; `segment_truncate` arms a 10 ms window.setTimeout which is run at the end, like
oracle/run_reference_modules.js does under Node."""
from __future__ import annotations

import os

from .interp import Interp, JSArray, JSObject, JSThrow, JSTyped, Native, UNDEF
from .parser import matching_end

BUNDLE = "/root/reference/dist/main.js"
_ANCHORS = {            # inner module id -> first export registered by the module (unique in the bundle)
    0: 'n.d(t,"solve_poly"',
    3: 'n.d(t,"reset_segmentation"',
    4: 'n.d(t,"formant_features"',
    7: 'n.d(t,"get_utterance_features"',
}
_HEAD = 'function(e,t,n){"use strict";n.r(t),'


def available() -> bool:
    return os.path.exists(BUNDLE)


def module_sources(bundle_path: str = BUNDLE) -> dict:
    """{inner module id: (byte offset, 'function(e,t,n){...}')} cut out of the reference bundle."""
    with open(bundle_path, encoding="utf-8") as f:
        src = f.read()
    base = src.index("\n") + 1 if src.startswith("/*") else 0        # offsets are quoted relative to line 2
    out = {}
    for mid, anchor in _ANCHORS.items():
        a = src.index(anchor)
        if src.find(anchor, a + 1) >= 0:
            raise RuntimeError(f"anchor for module {mid} is not unique")
        start = src.rindex(_HEAD, 0, a)
        if a - start != len(_HEAD):
            raise RuntimeError(f"module {mid}: unexpected prologue")
        brace = start + len("function(e,t,n)")
        end = matching_end(src, brace)
        out[mid] = (start - base, src[start:end])
    return out


def _to_py(v):
    if isinstance(v, (JSArray, JSTyped)):
        return [_to_py(x) for x in v.a]
    if v is UNDEF:
        return None
    if isinstance(v, JSObject):
        return {k: _to_py(x) for k, x in v.props.items()}
    return v


class ReferenceModules:
    """The reference's modules, instantiated once (module state is re-initialised by reset_segmentation, like the app)."""

    def __init__(self, bundle_path: str = BUNDLE):
        self.it = it = Interp()
        self.sources = module_sources(bundle_path)
        self.cache = {}

        def req(this, args):
            mid = int(args[0])
            if mid in self.cache:
                return self.cache[mid].props['exports']
            exports = JSObject()
            module = JSObject({'exports': exports})
            self.cache[mid] = module
            if mid == 5:
                # numeric@1.2.6 (level 12 only): the reference's own hand-written functions + natives for the generated
                # element-wise helpers (oracle/minijs/numeric_shim.py)
                from . import numeric_shim
                with open(bundle_path, encoding="utf-8") as f:
                    module.props['exports'] = numeric_shim.install(it, f.read())
                return module.props['exports']
            if mid not in self.sources:
                return exports
            fn = it.eval_expression('(' + self.sources[mid][1] + ')')
            it.call(fn, exports, [module, exports, self.require])
            return module.props['exports']

        def n_r(this, a):
            a[0].props['__esModule'] = True
            return UNDEF

        def n_d(this, a):
            obj, name, getter = a
            if obj.getters is None:
                obj.getters = {}
            obj.getters[name] = getter
            return UNDEF

        self.require = Native(req, 'require')
        self.require.props = {'r': Native(n_r, 'r'), 'd': Native(n_d, 'd')}
        self.seg = it.call(self.require, UNDEF, [3.0])

    def fn(self, name):
        return self.it.get(self.seg, name)

    def analyze(self, frames, level, step_ms, bands=None, plot_len=200, pause_ms=200.0, minlen_ms=50.0, auto_gate=True,
                max_db=100.0, min_db=10.0, test_play=False, labels=()):
        """frames: iterable of B uint32 -> dict(seg_ci, events, syl_ci, formants, console)."""
        it = self.it
        frames = [list(map(float, f)) for f in frames]
        bands = int(bands if bands is not None else len(frames[0]))
        events = []

        def cb(this, a):
            events.append([_to_py(x) for x in a])
            return UNDEF

        p = it.call(self.fn('reset_segmentation'), UNDEF,
                    [float(level), float(bands), float(plot_len), float(step_ms), float(pause_ms), float(minlen_ms),
                     bool(auto_gate), float(max_db), float(min_db), Native(cb, 'callback'), bool(test_play),
                     JSArray([float(x) for x in labels])])
        it.run_microtasks()
        # after a dropped segment callbacks_processed < seg_ci.length for ever and the reference polls
        # 20 x 500 ms ("await_busy_last_process timeout") before it resets: run those timers
        while p.state == 0 and it.timers:
            it.run_timers()
        if p.state != 1:
            raise RuntimeError("reset_segmentation did not resolve: " + repr(_to_py(p.value)))
        it.console.clear()
        push = self.fn('spectrum_push')
        for idx, f in enumerate(frames):
            it.call(push, UNDEF, [JSTyped('Uint32Array', f), float(idx)])
            it.run_microtasks()                    # one MessagePort message per frame
        it.call(self.fn('segment_truncate'), UNDEF, [])
        it.run_timers()
        it.run_microtasks()
        n = int(it.call(self.fn('get_segments_count'), UNDEF, [4.0]))          # stores (a dropped segment has none)
        seg_ci = []
        while True:                                                             # seg_ci keeps dropped segments too
            ci = it.call(self.fn('get_segments_ci'), UNDEF, [float(len(seg_ci))])
            if ci is UNDEF:
                break
            seg_ci.append([int(x) for x in _to_py(ci)])
        out = {
            "seg_ci": seg_ci,
            "stored": n,
            "events": events,
            "console": list(it.console),
        }
        if level >= 4:
            out["formants"] = [_to_py(it.call(self.fn('get_segment'), UNDEF, [float(e), 4.0])) for e in range(n)]
        if level in (10, 11, 12, 13):
            out["syl_ci"] = [[[int(v) for v in s] for s in _to_py(it.call(self.fn('get_syllables_ci'), UNDEF, [float(e)]))]
                             for e in range(n)]
        return out


__all__ = ["ReferenceModules", "available", "module_sources", "JSThrow"]


# ---------------------------------------------------------------------------------------------------------------
# The reference's PUBLIC API module (inner module 1 @B2714: configure / LaunchAudioNodes / StopAudioNodes /
# set_predicted_label_for_segment), executed with its audio-graph module (inner module 2, browser-only) replaced by a
# recording stub: what configure() leaves in the settings object, which string a LaunchAudioNodes call rejects with, and
# which calls it makes into module 2 in which order with which arguments.
_API_ANCHOR = 'n.d(t,"configure"'


class ReferenceAPI:
    def __init__(self, bundle_path: str = BUNDLE):
        with open(bundle_path, encoding="utf-8") as f:
            src = f.read()
        a = src.index(_API_ANCHOR)
        start = src.rindex(_HEAD, 0, a)
        end = matching_end(src, start + len("function(e,t,n)"))
        self.source = src[start:end]
        self.it = it = Interp()
        self.calls = []
        self.playing = False
        self.fail = {}          # module-2 function name -> rejection value

        def stub(name, ret_promise=True):
            def f(this, args):
                self.calls.append((name, [_to_py(x) for x in args]))
                if not ret_promise:
                    return UNDEF
                p = it.eval_expression("new Promise(function(a,b){window.__r=a;window.__j=b})")
                w = it.globals.vars['window']
                if name in self.fail:
                    it.call(it.get(w, '__j'), UNDEF, [self.fail[name]])
                else:
                    it.call(it.get(w, '__r'), UNDEF, [True])
                return p
            return Native(f, name)

        mod2 = JSObject({n: stub(n) for n in ("reset_nodes", "reset_segmentor", "reset_plot", "offline_play_the_file",
                                                "online_play_the_file", "online_play_the_sop", "online_play_the_mic")})
        mod2.props["isNodePlaying"] = Native(lambda this, args: self.playing, "isNodePlaying")
        for n in ("Garbage_Collect", "disconnect_nodes", "set_predicted_label_for_segment"):
            mod2.props[n] = stub(n, ret_promise=False)
        exports = JSObject()
        module = JSObject({'exports': exports})

        def req(this, args):
            assert int(args[0]) == 2
            return mod2

        def n_r(this, a_):
            return UNDEF

        def n_d(this, a_):
            if a_[0].getters is None:
                a_[0].getters = {}
            a_[0].getters[a_[1]] = a_[2]
            return UNDEF

        require = Native(req, 'require')
        require.props = {'r': Native(n_r, 'r'), 'd': Native(n_d, 'd')}
        fn = it.eval_expression('(' + self.source + ')')
        it.call(fn, exports, [module, exports, require])
        self.exports = exports

    def _js(self, v):
        if isinstance(v, dict):
            return JSObject({k: self._js(x) for k, x in v.items()})
        if isinstance(v, (list, tuple)):
            return JSArray([self._js(x) for x in v])
        if isinstance(v, bool) or v is None or isinstance(v, str):
            return v
        return float(v)

    def settings(self) -> dict:
        """The module-level settings object (`a` @B2972) as configure() left it."""
        env = self.it.get(self.exports, 'configure').env
        return _to_py(env.vars['a'])

    def configure(self, cfg: dict) -> dict:
        self.it.call(self.it.get(self.exports, 'configure'), UNDEF, [self._js(cfg)])
        return self.settings()

    def launch(self, *args):
        """-> ('resolved', value) | ('rejected', value), plus self.calls = the calls made into the audio-graph module."""
        self.calls.clear()
        p = self.it.call(self.it.get(self.exports, 'LaunchAudioNodes'), UNDEF, [self._js(a) if not isinstance(a, Native) else a for a in args])
        self.it.run_microtasks()
        return ('resolved' if p.state == 1 else 'rejected' if p.state == 2 else 'pending', _to_py(p.value))
