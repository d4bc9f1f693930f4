#!/usr/bin/env node
'use strict';
/*
 * TEST INFRASTRUCTURE ONLY.  Runs the REFERENCE'S OWN minified segmentor / formants / stats modules on uint32 frames.
 *
 *   node oracle/run_reference_modules.js <path/to/dist/main.js> <frames.json> [level] [step_ms]
 *
 * frames.json = {"bands": B, "frames": [[...B uint32...], ...]} (e.g. dumped from oracle.frontend()).  Evaluates
 * webpack module 584 (formantanalyzer@1.1.6, byte range @B100-B114174 of dist/main.js line 2) with a tiny
 * webpack-require shim, pushes every frame through spectrum_push, calls segment_truncate, and prints the callbacks
 * as JSON.  Comparing this output with oracle/fa_oracle.c upgrades the oracle from "restatement" to "checked against
 * the reference's own code".  It cannot run in the build image (no JS engine); use it wherever `node` exists with a
 * checkout of tabahi/WebSpeechAnalyzer.
 */
const fs = require('fs');
const [, , bundlePath, framesPath, levelArg, stepArg] = process.argv;
if (!bundlePath || !framesPath) { console.error('usage: run_reference_modules.js dist/main.js frames.json [level] [step_ms]'); process.exit(2); }
const src = fs.readFileSync(bundlePath, 'utf8');
// module 584 = `584:function(module){var n;n=function(){return function(e){ <webpack runtime> n(n.s=1)}([m0, m1, ...])},
// module.exports=n()}`: take the inner module array [m0 .. m8] and drive it with our own require shim.
const start = src.indexOf('584:function(module)');
if (start < 0) { console.error('module 584 (formantanalyzer) not found in the bundle'); process.exit(3); }
global.window = { setTimeout: (f) => f(), AudioContext: function () {} };
global.self = global.window;
const innerStart = src.indexOf('([function(e,t,n){"use strict";', start);
const innerEnd = src.indexOf('},module.exports=n()', innerStart);
if (innerStart < 0 || innerEnd < 0) { console.error('inner module array not found'); process.exit(3); }
const mods = eval(src.slice(innerStart + 1, innerEnd - 1));     // '[' ... ']'
const cache = {};
function req(id) {
  if (cache[id]) return cache[id].exports;
  const m = cache[id] = { exports: {} };
  mods[id].call(m.exports, m, m.exports, req);
  return m.exports;
}
req.r = (e) => Object.defineProperty(e, '__esModule', { value: true });
req.d = (e, name, getter) => Object.defineProperty(e, name, { enumerable: true, get: getter });
req.n = (e) => { const g = e && e.__esModule ? () => e.default : () => e; req.d(g, 'a', g); return g; };
req.o = (o, p) => Object.prototype.hasOwnProperty.call(o, p);
const seg = req(3);                          // inner module 3: segmentor (@B23403)
const input = JSON.parse(fs.readFileSync(framesPath, 'utf8'));
const level = parseInt(levelArg || '13', 10), step = parseFloat(stepArg || '15');
const events = [];
seg.reset_segmentation(level, input.bands, 200, step, 200, 50, true, 100, 10, (...args) => events.push(args), false, []).then(() => {
  input.frames.forEach((f, idx) => seg.spectrum_push(Uint32Array.from(f), idx));
  seg.segment_truncate();
  setTimeout(() => {
    const n = seg.get_segments_count(4);
    const ci = [];
    for (let e = 0; e < n; e++) ci.push(seg.get_segments_ci(e));
    console.log(JSON.stringify({ seg_ci: ci, events: events.map((e) => [e[0], e[2], Array.from(e[3] || [], (r) => Array.from(r))]) }));
  }, 50);
});
