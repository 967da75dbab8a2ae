// main_multi.mjs -- the Node host on all GPUs of one box: one child process per GPU, frame partitioned by tiles x sample
// sets exactly like fspt_b200/dist.py (the tested Python twin), collectives inside libfspt_b200.so (NCCL over NVLink).
// Not runnable in this image (no Node).
//
//   node main_multi.mjs scene.flat.json out.rgba 3840 2160 1024 8        (last argument: number of GPUs)
//
// The parent creates the NCCL unique id and hands it to the children through their environment; rank 0 is the only
// process that reads the scene file: it uploads once (fspt_scene_upload) and every other GPU receives the device-
// resident records with fspt_scene_broadcast.  Each rank renders its rectangle / tick subset into its own f32 sum
// buffer; one fspt_reduce_accum per frame; rank 0 resolves (draw.fs) and writes the RGBA8 frame.
import { createRequire } from 'module';
import { spawn } from 'child_process';
import fs from 'fs';
const require = createRequire(import.meta.url);
const fspt = require('./build/Release/fspt_napi.node');

function mulberry32(a) {
  return () => { a |= 0; a = a + 0x6D2B79F5 | 0; let t = Math.imul(a ^ a >>> 15, 1 | a);
    t = t + Math.imul(t ^ t >>> 7, 61 | t) ^ t; return ((t ^ t >>> 14) >>> 0) / 4294967296; };
}
const b64 = (s, T) => { const b = Buffer.from(s, 'base64'); return new T(b.buffer, b.byteOffset, b.length / T.BYTES_PER_ELEMENT); };

// fspt_b200.dist.tile_grid / partition
function tileGrid(world, w, h) {
  let nTiles = 1;
  while (nTiles * 2 <= world && world % (nTiles * 2) === 0 && Math.floor(w * h / nTiles) * 64 > (64 << 20)) nTiles *= 2;
  return [nTiles, world / nTiles];
}
function partition(rank, world, w, h, nSamples) {
  const [nTiles, nSets] = tileGrid(world, w, h);
  const tile = Math.floor(rank / nSets), sset = rank % nSets, rows4 = Math.floor((h + 3) / 4);
  const r0 = Math.floor(rows4 * tile / nTiles) * 4;
  const r1 = tile === nTiles - 1 ? h : Math.min(h, Math.floor(rows4 * (tile + 1) / nTiles) * 4);
  const ticks = [];
  for (let k = sset; k < nSamples; k += nSets) ticks.push(k);
  return { rect: [0, r0, w, r1 - r0], ticks };
}

const [, , scenePath, outPath, W = '1280', H = '720', SPP = '64', G = '1'] = process.argv;
const world = parseInt(G), resolution = [parseInt(W), parseInt(H)], spp = parseInt(SPP);

if (process.env.FSPT_RANK === undefined) {           // parent: spawn one process per GPU
  const id = Buffer.from(fspt.commUniqueId()).toString('base64');
  const kids = [];
  for (let r = 0; r < world; r++) {
    kids.push(new Promise((res, rej) => spawn(process.execPath, process.argv.slice(1),
      { stdio: 'inherit', env: { ...process.env, FSPT_RANK: String(r), FSPT_NCCL_ID: id } })
      .on('exit', c => c === 0 ? res() : rej(new Error('rank ' + r + ' exited with ' + c)))));
  }
  await Promise.all(kids);
} else {                                             // child: rank r on GPU r
  const rank = parseInt(process.env.FSPT_RANK);
  const ctx = fspt.create(resolution[0], resolution[1], rank);
  fspt.commInit(ctx, new Uint8Array(Buffer.from(process.env.FSPT_NCCL_ID, 'base64')), rank, world);
  let s = null;
  if (rank === 0) {
    s = JSON.parse(fs.readFileSync(scenePath, 'utf8'));
    fspt.sceneUpload(ctx, {
      bvh: b64(s.bvh, Float32Array), triangles: b64(s.triangles, Float32Array), materials: b64(s.materials, Float32Array),
      normals: b64(s.normals, Float32Array), uvs: b64(s.uvs, Float32Array), atlas: b64(s.atlas, Uint8Array),
      env: b64(s.env, Uint8Array), radianceBins: b64(s.radianceBins, Uint16Array),
      atlasRes: s.atlasRes, atlasLayers: s.atlasLayers, envWidth: s.envWidth, envHeight: s.envHeight, leafSize: 4,
    });
  }
  fspt.sceneBroadcast(ctx, 0);
  const cam = JSON.parse(fs.readFileSync(scenePath + '.camera.json', 'utf8'));   // small sidecar every rank may read
  const frame = { eye: cam.cameraPos, dir: cam.cameraDir, fovScale: cam.fovScale || 0.5,
    lensFeatures: [1 - 1 / (cam.focalDepth || 2.0), cam.aperture ?? 0.02], envTheta: cam.environmentTheta || 0 };
  const rnd = mulberry32(1), rcAll = new Float32Array(spp), rtAll = new Float32Array(spp);
  for (let i = 0; i < spp; i++) { rcAll[i] = rnd() * 10000; rtAll[i] = rnd() * 10000; }
  const { rect, ticks } = partition(rank, world, resolution[0], resolution[1], spp);
  fspt.setAccumMode(ctx, 1);                         // f32 sum + per-pixel sample count
  fspt.setTile(ctx, ...rect);
  fspt.clear(ctx);
  if (ticks.length) fspt.render(ctx, frame, 0, Float32Array.from(ticks, k => rcAll[k]), Float32Array.from(ticks, k => rtAll[k]));
  fspt.reduceAccum(ctx, 0);                          // ncclReduce(sum) over NVLink, on the library's stream
  if (rank === 0) {
    const rgba = fspt.resolve(ctx, { exposure: cam.exposure || 1, saturation: 1, maxSigma: 2, scale: 1, denoise: 0 },
      resolution[0], resolution[1]);
    fs.writeFileSync(outPath, Buffer.from(rgba.buffer));
  } else {
    fspt.readAccum(ctx, resolution[0], resolution[1]);   // synchronise before the process exits
  }
}
