// inject.js -- loaded as a classic script in <head>, BEFORE the reference's ES modules, by the harness page that
// oracle/swiftshader/run_harness.py serves.  It does not modify any reference file; it only
//   1. replaces the host RNG: main.js draws `Math.random()*10000` twice per pass (main.js:748,777); here it is
//      mulberry32(seed) -- the same stream as fspt_b200.scenes.rand_bases -- and every value handed out is logged;
//   2. captures the WebGL2 context of the `trace` canvas (main.js:78) and watches drawArrays: per tick() the
//      reference issues drawCamera (:755), drawTracer (:806), drawQuad (:823).  After the camera draw the two
//      RGBA32F camera targets are read back (readBuffer COLOR_ATTACHMENT0 / 1, FLOAT readPixels), after the tracer
//      draw the accumulation target (with mode=test that is the patched bvh_test.fs tail: index, t, count);
//   3. removes the vsync cap for timing runs (requestAnimationFrame -> immediate callback, main.js:855) and brackets
//      every tick with gl.finish();
//   4. POSTs the arrays to /result on the harness server.
// Query parameters: seed, ticks (passes to capture after the reference's initial clear()), timing=1.
(function () {
  'use strict';
  const Q = new URLSearchParams(window.location.search);
  const SEED = parseInt(Q.get('seed') || '1', 10) >>> 0;
  const TICKS = parseInt(Q.get('ticks') || '2', 10);
  const TIMING = Q.get('timing') === '1';

  let state = SEED;
  const randLog = [];
  Math.random = function () {  // mulberry32
    state = (state + 0x6D2B79F5) >>> 0;
    let t = state;
    t = Math.imul(t ^ (t >>> 15), t | 1);
    t ^= t + Math.imul(t ^ (t >>> 7), t | 61);
    const v = ((t ^ (t >>> 14)) >>> 0) / 4294967296;
    randLog.push(v);
    return v;
  };

  if (TIMING) window.requestAnimationFrame = function (cb) { return setTimeout(cb, 0); };

  const captured = { passes: [], tickMs: [], w: 0, h: 0 };
  let gl = null, draws = 0, cleared = false, tickStart = 0, done = false;

  function post(path, payload) {
    return fetch(path, { method: 'POST', body: payload });
  }
  function readFloat(g, attachment) {
    const w = g.drawingBufferWidth, h = g.drawingBufferHeight;
    const fbo = g.getParameter(g.DRAW_FRAMEBUFFER_BINDING);
    g.bindFramebuffer(g.READ_FRAMEBUFFER, fbo);
    g.readBuffer(attachment);
    const px = new Float32Array(w * h * 4);
    g.readPixels(0, 0, w, h, g.RGBA, g.FLOAT, px);
    return px;
  }
  function finish() {
    if (done) return;
    done = true;
    const w = captured.w, h = captured.h;
    const header = JSON.stringify({ width: w, height: h, passes: captured.passes.length, rand: randLog,
                                    tick_ms: captured.tickMs, user_agent: navigator.userAgent,
                                    renderer: gl.getParameter(gl.RENDERER) });
    const hb = new TextEncoder().encode(header);
    const n = captured.passes.length;
    const body = new Uint8Array(8 + hb.length + n * 3 * w * h * 16);
    new DataView(body.buffer).setUint32(0, hb.length, true);
    new DataView(body.buffer).setUint32(4, n, true);
    body.set(hb, 8);
    let off = 8 + hb.length;
    for (const p of captured.passes) {
      for (const a of [p.pos, p.dir, p.accum]) { body.set(new Uint8Array(a.buffer), off); off += a.byteLength; }
    }
    post('/result', body).then(() => { document.title = 'FSPT_HARNESS_DONE'; });
  }

  const origGetContext = HTMLCanvasElement.prototype.getContext;
  HTMLCanvasElement.prototype.getContext = function (kind, attrs) {
    const ctx = origGetContext.call(this, kind, attrs);
    if (kind !== 'webgl2' || this.id !== 'trace' || !ctx || gl) return ctx;
    gl = ctx;
    const origDraw = gl.drawArrays.bind(gl), origClear = gl.clear.bind(gl);
    let current = null;
    // main.js clear() (:826-836) runs once after the first tick (`dirty` starts true) and discards that pass:
    // capture starts with the pass after it
    gl.clear = function (mask) { const r = origClear(mask); cleared = true; draws = 0; current = null; return r; };
    gl.drawArrays = function (mode, first, count) {
      const phase = draws % 3;  // 0 camera, 1 tracer, 2 quad
      if (phase === 0 && TIMING) { gl.finish(); tickStart = performance.now(); }
      const r = origDraw(mode, first, count);
      draws++;
      if (!cleared || done) return r;
      if (TIMING) {
        if (phase === 1) { gl.finish(); captured.tickMs.push(performance.now() - tickStart); }
        if (phase === 2 && captured.tickMs.length >= TICKS) { captured.w = gl.drawingBufferWidth; captured.h = gl.drawingBufferHeight; finish(); }
        return r;
      }
      if (phase === 0) {
        current = { pos: readFloat(gl, gl.COLOR_ATTACHMENT0), dir: readFloat(gl, gl.COLOR_ATTACHMENT1), accum: null };
      } else if (phase === 1 && current) {
        current.accum = readFloat(gl, gl.COLOR_ATTACHMENT0);
        captured.passes.push(current);
        current = null;
        captured.w = gl.drawingBufferWidth; captured.h = gl.drawingBufferHeight;
        if (captured.passes.length >= TICKS) finish();
      }
      return r;
    };
    return ctx;
  };
  window.addEventListener('error', function (e) { post('/error', String(e.message || e)); });
  window.addEventListener('unhandledrejection', function (e) { post('/error', String(e.reason)); });
})();
