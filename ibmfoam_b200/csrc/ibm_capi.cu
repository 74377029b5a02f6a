// C ABI implementation: context lifetime, shape-class tables, layout/transfer plumbing, loops 1-4 drivers,
// accessors and parity taps.  Loop 5 lives in ibm_motion.cu, NCCL in ibm_comm.cu.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <condition_variable>
#include <mutex>
#include <thread>

#include "ibm_ctx.h"

using namespace ibm;

static std::string g_createErr;

namespace ibm {

int ctx_fail(ibm_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg; else g_createErr = msg;
  return code;
}
bool ctx_cuda_ok(ibm_ctx* c, cudaError_t e, const char* what) {
  if (e == cudaSuccess) return true;
  std::string m = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
  if (c) c->err = m; else g_createErr = m;
  return false;
}

static cudaEvent_t get_event(ibm_ctx* c) {
  if (!c->eventPool.empty()) { cudaEvent_t e = c->eventPool.back(); c->eventPool.pop_back(); return e; }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
static void resolve_events(ibm_ctx* c) {
  if (c->events.empty()) return;
  cudaStreamSynchronize(c->stream);
  for (auto& p : c->events) {
    float ms = 0.f;
    if (p.b && cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) c->accMs[p.cat] += ms;
    c->eventPool.push_back(p.a);
    if (p.b) c->eventPool.push_back(p.b);
  }
  c->events.clear();
}
static void marks_resolve(ibm_ctx* c) {
  if (c->dbgMarks.empty()) return;
  cudaStreamSynchronize(c->stream);
  auto add = [](std::vector<std::pair<std::string, std::vector<float>>>& acc, const std::string& k, double v) {
    for (auto& e : acc) if (e.first == k) { e.second.push_back((float)v); return; }
    acc.push_back({k, {(float)v}});
  };
  for (size_t i = 1; i < c->dbgMarks.size(); ++i) {
    float ms = 0;
    cudaEventElapsedTime(&ms, c->dbgMarks[i - 1].second, c->dbgMarks[i].second);
    const std::string k = std::string(c->dbgMarks[i - 1].first) + " -> " + c->dbgMarks[i].first;
    add(c->dbgAcc, k, ms);
    add(c->dbgAccHost, k, (c->dbgHost[i] - c->dbgHost[i - 1]) * 1e3);
  }
  for (auto& m : c->dbgMarks) cudaEventDestroy(m.second);
  c->dbgMarks.clear(); c->dbgHost.clear();
}
void ctx_mark(ibm_ctx* c, const char* label) {
  if (!c->dbgTiming) return;
  if (c->dbgMarks.size() >= 512) marks_resolve(c);
  cudaEvent_t e;
  cudaEventCreate(&e);
  cudaEventRecord(e, c->stream);
  c->dbgMarks.push_back({label, e});
  timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
  c->dbgHost.push_back(ts.tv_sec + 1e-9 * ts.tv_nsec);
}
void ctx_marks_report(ibm_ctx* c) {
  if (!c->dbgTiming) return;
  marks_resolve(c);
  int rank = comm_rank(c);
  auto median = [](std::vector<float> v) { std::sort(v.begin(), v.end()); return v.empty() ? 0.0f : v[v.size() / 2]; };
  for (size_t i = 0; i < c->dbgAcc.size(); ++i)
    fprintf(stderr, "[timing rank %d] %-44s median: device %8.3f ms  host %8.3f ms  (n=%zu)\n", rank, c->dbgAcc[i].first.c_str(),
            median(c->dbgAcc[i].second), median(c->dbgAccHost[i].second), c->dbgAcc[i].second.size());
  for (int q = 2; q <= 8; ++q) {
    long long v = 0;
    char name[32];
    snprintf(name, sizeof(name), "loop5_stat%d", q);
    if (motion_counter(c, name, &v)) fprintf(stderr, "[timing rank %d] %s = %lld\n", rank, name, v);
  }
}
void ctx_tic(ibm_ctx* c, int cat) {
  if (c->events.size() >= 2048) resolve_events(c);
  EventPair p{get_event(c), nullptr, cat};
  cudaEventRecord(p.a, c->stream);
  c->events.push_back(p);
}
void ctx_toc(ibm_ctx* c) {
  EventPair& p = c->events.back();
  p.b = get_event(c);
  cudaEventRecord(p.b, c->stream);
}

int ctx_ensure_stage(ibm_ctx* c, size_t doubles) {
  if (c->stageDoubles >= doubles) return IBM_OK;
  if (c->dStage) cudaFree(c->dStage);
  c->dStage = nullptr; c->stageDoubles = 0;
  IBM_CUDA(c, cudaMalloc(&c->dStage, doubles * sizeof(double)));
  c->stageDoubles = doubles;
  return IBM_OK;
}
int ctx_ensure_pinned(ibm_ctx* c, size_t doubles) {
  if (c->hPinnedDoubles >= doubles) return IBM_OK;
  if (c->hPinned) cudaFreeHost(c->hPinned);
  c->hPinned = nullptr; c->hPinnedDoubles = 0;
  IBM_CUDA(c, cudaMallocHost(&c->hPinned, doubles * sizeof(double)));
  c->hPinnedDoubles = doubles;
  return IBM_OK;
}

static int ensure_field(ibm_ctx* c, double** p, int nComp) {
  if (*p) return IBM_OK;
  IBM_CUDA(c, cudaMalloc(p, (size_t)nComp * c->L.nCells * sizeof(double)));
  return IBM_OK;
}

// ---- shape classes: centeredMesh_ / Sf_ exactly as discretize()/calcSf() evaluate them (host libm) ----
static void build_class(ShapeClassHost& h, int kind, double d, double a, double b, int nTheta, double delta, double l) {
  const double Pi = IBM_PI;
  ShapeClassDev& s = h.dev;
  s.kind = kind; s.nTheta = nTheta; s.d = d; s.a = a; s.b = b; s.delta = delta; s.l = l;
  s.nk = (kind == IBM_SPHERE) ? nTheta / 2 : 1;  // sphere.C:66-69 (always), cylinder.C:72, ellipse.C:89
  s.N = 3 * nTheta * s.nk;
  const int nk = s.nk;
  h.centered.assign((size_t)3 * s.N, 0.0);
  h.Sf.assign((size_t)3 * nTheta * nk, 0.0);
  h.cosT.resize(nTheta); h.sinT.resize(nTheta); h.sinP.assign(nk, 0.0); h.cosP.assign(nk, 1.0);
  h.Rtab.assign((size_t)3 * nTheta, 0.0);
  h.ck.assign(nk, 1.0);
  s.dTheta = 2.0 * Pi / (nTheta);
  const double dl[3] = {-1 * delta, 0, delta};
  const double dTheta = 2.0 * Pi / (nTheta);
  auto idx = [&](int i, int j, int k) { return i + j * 3 + k * 3 * nTheta; };
  auto idx2 = [&](int i, int k) { return i + k * nTheta; };
  for (int j = 0; j < nTheta; ++j) { double th = double(j) * dTheta; h.cosT[j] = std::cos(th); h.sinT[j] = std::sin(th); }
  if (kind == IBM_SPHERE) {
    const double R = d / 2.0;
    const double dPhi = Pi / (nk - 1);
    for (int k = 0; k < nk; ++k) { double phi = dPhi * double(k); h.sinP[k] = std::sin(phi); h.cosP[k] = std::cos(phi); }
    for (int i = 0; i < 3; ++i) {
      double r = R + dl[i];
      for (int j = 0; j < nTheta; ++j) {
        h.Rtab[(size_t)i * nTheta + j] = r;
        double th = double(j) * dTheta;
        for (int k = 0; k < nk; ++k) {
          double phi = dPhi * double(k);
          double* p = &h.centered[(size_t)3 * idx(i, j, k)];
          p[0] = r * std::cos(th) * std::sin(phi);  // sphere.C:184-190
          p[1] = r * std::sin(th) * std::sin(phi);
          p[2] = r * std::cos(phi);
        }
      }
    }
    const double r = d / 2.0;
    for (int i = 0; i < nTheta; ++i) {
      double th = dTheta * double(i);
      for (int k = 0; k < nk; ++k) {
        double phi = dPhi * double(k);
        double f = (r * r) * dTheta;
        (void)f;
        double v[3] = {-(std::cos(th) * std::sin(phi)), -(std::sin(th) * std::sin(phi)), -(std::cos(phi))};
        for (int q = 0; q < 3; ++q) v[q] = v[q] * (r * r) * dTheta;  // -vector(...)*sqr(r)*dTheta  (sphere.C:134-140)
        double ck;
        if (k == 0) ck = (std::cos(phi) - std::cos(phi + dPhi / 2.0));
        else if (k == nk - 1) ck = (std::cos(phi - dPhi / 2.0) - std::cos(phi));
        else ck = (std::cos(phi - dPhi / 2.0) - std::cos(phi + dPhi / 2.0));
        for (int q = 0; q < 3; ++q) h.Sf[(size_t)3 * idx2(i, k) + q] = v[q] * ck;
        h.ck[k] = ck;
      }
    }
    s.Rbound = d / 2.0;
    s.V = 1.0 / 6.0 * Pi * (d * d * d);        // sphere.C:262-265
    s.I = 2.0 / 5.0 * ((d / 2.0) * (d / 2.0));  // sphere.C:64
    s.A = Pi * ((d / 2.0) * (d / 2.0));        // sphere.C:256-259
  } else if (kind == IBM_CYLINDER) {
    const double R = d / 2.0;
    for (int i = 0; i < 3; ++i) {
      double r = R + dl[i];
      for (int j = 0; j < nTheta; ++j) {
        h.Rtab[(size_t)i * nTheta + j] = r;
        double th = double(j) * dTheta;
        double* p = &h.centered[(size_t)3 * idx(i, j, 0)];
        p[0] = r * std::cos(th); p[1] = r * std::sin(th); p[2] = 0.0;  // cylinder.C:142-148
      }
    }
    const double r = d / 2.0;
    for (int i = 0; i < nTheta; ++i) {
      double th = dTheta * double(i);
      double v[3] = {-std::cos(th), -std::sin(th), -0.0};
      for (int q = 0; q < 3; ++q) h.Sf[(size_t)3 * idx2(i, 0) + q] = v[q] * r * dTheta * l;  // cylinder.C:116-122
    }
    s.Rbound = d / 2.0;
    s.V = Pi * ((d / 2.0) * (d / 2.0)) * l;  // cylinder.C:215-220
    s.I = (d / 8.0) * (d / 8.0);             // cylinder.C:71 (as coded)
    s.A = d * l;                             // cylinder.C:210-213
  } else {
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < nTheta; ++j) {
        double th = double(j) * dTheta;
        double ca = std::cos(th) / (a), sb = std::sin(th) / (b);
        double R = std::sqrt(1.0 / (ca * ca + sb * sb)) + dl[i];  // ellipse.C:135-137
        h.Rtab[(size_t)i * nTheta + j] = R;
        double* p = &h.centered[(size_t)3 * idx(i, j, 0)];
        p[0] = R * std::cos(th); p[1] = R * std::sin(th); p[2] = 0.0;
      }
    for (int i = 0; i < nTheta; ++i)
      for (int q = 0; q < 3; ++q)
        h.Sf[(size_t)3 * idx2(i, 0) + q] = -h.centered[(size_t)3 * idx(1, i, 0) + q] * dTheta * l;  // ellipse.C:57
    s.Rbound = std::fmax(a, b);
    s.V = Pi * a * b * l;     // ellipse.C:234-237
    s.I = (a * a) + (b * b);  // ellipse.C:88 (as coded)
    s.A = a * b * Pi;         // ellipse.C:229-232
  }
}

static int upload_classes(ibm_ctx* c) {
  for (auto p : c->classPools) cudaFree(p);
  c->classPools.clear();
  std::vector<ShapeClassDev> devs;
  for (auto& h : c->classes) {
    size_t n = h.centered.size() + h.Sf.size() + h.cosT.size() + h.sinT.size() + h.sinP.size() + h.cosP.size() + h.Rtab.size() + h.ck.size();
    double* pool = nullptr;
    IBM_CUDA(c, cudaMalloc(&pool, n * sizeof(double)));
    c->classPools.push_back(pool);
    std::vector<double> flat;
    flat.reserve(n);
    auto put = [&](const std::vector<double>& v, const double*& dst) {
      dst = pool + flat.size();
      flat.insert(flat.end(), v.begin(), v.end());
    };
    put(h.centered, h.dev.centered); put(h.Sf, h.dev.Sf); put(h.cosT, h.dev.cosT); put(h.sinT, h.dev.sinT);
    put(h.sinP, h.dev.sinP); put(h.cosP, h.dev.cosP); put(h.Rtab, h.dev.Rtab); put(h.ck, h.dev.ck);
    IBM_CUDA(c, cudaMemcpy(pool, flat.data(), n * sizeof(double), cudaMemcpyHostToDevice));
    devs.push_back(h.dev);
  }
  if (c->dClasses) cudaFree(c->dClasses);
  c->dClasses = nullptr;
  if (!devs.empty()) {
    IBM_CUDA(c, cudaMalloc(&c->dClasses, devs.size() * sizeof(ShapeClassDev)));
    IBM_CUDA(c, cudaMemcpy(c->dClasses, devs.data(), devs.size() * sizeof(ShapeClassDev), cudaMemcpyHostToDevice));
  }
  return IBM_OK;
}

static int alloc_instances(ibm_ctx* c, InstanceSoA& s, int cap) {
  auto ai = [&](int** p) -> bool {  // (one spare slot: ctx_update sends a flag behind the last instance's point count)
    if (!ctx_cuda_ok(c, cudaMalloc(p, ((size_t)cap + 1) * sizeof(int)), "cudaMalloc(instances)")) return false;
    cudaMemset(*p, 0, ((size_t)cap + 1) * sizeof(int));
    c->instAllocs.push_back(*p);
    return true;
  };
  auto ad = [&](double** p, int mult = 1) -> bool {
    if (!ctx_cuda_ok(c, cudaMalloc(p, (size_t)cap * mult * sizeof(double)), "cudaMalloc(instances)")) return false;
    cudaMemset(*p, 0, (size_t)cap * mult * sizeof(double));
    c->instAllocs.push_back(*p);
    return true;
  };
  bool ok = ai(&s.index) && ai(&s.copy) && ai(&s.active) && ai(&s.onMesh) && ai(&s.cls) && ai(&s.nShell) && ai(&s.shellOff) &&
            ai(&s.nPtsOnMesh) && ai(&s.onLocal) && ad(&s.rho) && ad(&s.age) && ad(&s.rot, 9);
  for (int q = 0; q < 3 && ok; ++q) ok = ad(&s.c[q]) && ad(&s.th[q]) && ad(&s.v[q]) && ad(&s.om[q]) && ad(&s.F[q]) && ad(&s.T[q]);
  return ok ? IBM_OK : IBM_E_CUDA;
}

static int clear_live_F(ibm_ctx* c) {
  if (c->fLive && c->dF) {
    launch_clear_shell_cells(c->stream, c->dShellCell, c->totalShell, (size_t)c->L.nCells, c->dF);
    c->nLaunch++;
  }
  c->fLive = false;
  return IBM_OK;
}

// particleIBM::update() for all instances + deleteInActiveParticles + updateI
int ctx_update(ibm_ctx* c) {
  clear_live_F(c);
  ctx_tic(c, T_UPDATE);
  if (c->shellCap == 0) {
    // first guess: volume of the shell / cell volume per instance, with slack; grown on overflow
    size_t guess = 1024;
    const double cellV = c->L.dx * c->L.dy * (c->nGeometricD == 2 ? 1.0 : c->L.dz);
    for (int i = 0; i < c->nParticles; ++i) {
      const ShapeClassDev& s = c->classes[c->origCls[i]].dev;
      const double R = s.Rbound, r = std::max(0.0, R - s.delta);
      const double vol = c->nGeometricD == 2 ? IBM_PI * (R * R - r * r) : 4.0 / 3.0 * IBM_PI * (R * R * R - r * r * r);
      guess += (size_t)(1.3 * vol / cellV) + 16;
    }
    c->shellCap = (int)std::min<size_t>(guess * 3 / 2, 0x7fffffff);
    IBM_CUDA(c, cudaMalloc(&c->dShellCell, (size_t)c->shellCap * sizeof(int)));
    IBM_CUDA(c, cudaMalloc(&c->dShellJK, (size_t)c->shellCap * sizeof(int)));
  }
  c->sched.chunk = (c->forcingVariant == 2 && c->chunk256) ? 256 : FORCING_CHUNK;
  for (int attempt = 0; attempt < 8; ++attempt) {
    IBM_CUDA(c, cudaMemsetAsync(c->dScal, 0, 8 * sizeof(int), c->stream));
    // dScal: [0] kept instances, [1] shell cursor, [2] work items, [3] updates per forcing call, [4] shell overflow
    ctx_mark(c, "update:begin");
    // The Morton order of the schedule and the geometry of the launch records need the centres only: they run on a side
    // stream beside the marking kernel (0.5 ms of latency-bound kernels off the critical path); what depends on the marking
    // (work-item table, shell segments, has-local-points flags) follows on the main stream after the join.
    if (!c->updSide) {
      IBM_CUDA(c, cudaStreamCreateWithFlags(&c->updSide, cudaStreamNonBlocking));
      IBM_CUDA(c, cudaEventCreateWithFlags(&c->updFork, cudaEventDisableTiming));
      IBM_CUDA(c, cudaEventCreateWithFlags(&c->updJoin, cudaEventDisableTiming));
    }
    IBM_CUDA(c, cudaEventRecord(c->updFork, c->stream));
    IBM_CUDA(c, cudaStreamWaitEvent(c->updSide, c->updFork, 0));
    if (c->nInst > 0) {
      launch_schedule_order(c->updSide, c->L, c->in, c->nInst, c->sched);
      if (c->v7ok) { launch_v7_prepare(c->updSide, c->L, c->dClasses, c->in, c->nInst, c->sched, c->v7capF, c->v7prep); c->nLaunch++; }
    }
    IBM_CUDA(c, cudaEventRecord(c->updJoin, c->updSide));
    const bool nearOnly = comm_ranks(c) > 1 && c->markVariant == 1 && c->nInst > 0;
    if (nearOnly) {
      if (!c->dNearFlag) {
        IBM_CUDA(c, cudaMalloc(&c->dNearFlag, ((size_t)c->cap + 1) * sizeof(int)));
        IBM_CUDA(c, cudaMalloc(&c->dNearOff, ((size_t)c->cap + 1) * sizeof(int)));
        IBM_CUDA(c, cudaMalloc(&c->dNearList, ((size_t)c->cap + 1) * sizeof(int)));
      }
      launch_k1_far(c->stream, c->L, c->dClasses, c->in, c->nInst, c->sched, c->dNearFlag, c->dNearOff, c->dNearList, c->dScal + 6);
      c->nLaunch += 4;
    }
    launch_k1_mark(c->stream, c->L, c->dClasses, c->in, c->nInst, c->dShellCell, c->dShellJK, c->shellCap, c->dScal + 1, c->dScal + 4, c->markVariant,
                   nearOnly ? c->dNearList : nullptr, nearOnly ? c->dScal + 6 : nullptr);
    ctx_mark(c, "update:marked");
    if (comm_ranks(c) > 1 && c->nInst > 0) {
      // the on-mesh status of an instance is a property of the whole mesh: sum the point counts over the slabs.  The shell
      // overflow flag of this rank travels in the slot behind the last instance, so that every rank sees an overflow
      // anywhere and all of them take the same path (retry together instead of one rank leaving the collectives)
      launch_copy_int(c->stream, c->in.nPtsOnMesh + c->nInst, c->dScal + 4);
      int rcc = comm_allreduce_sum(c, c->in.nPtsOnMesh, (size_t)c->nInst + 1, 1);
      if (rcc) return rcc;
      launch_copy_int(c->stream, c->dScal + 4, c->in.nPtsOnMesh + c->nInst);
      c->nLaunch += 2;
    }
    ctx_mark(c, "update:counts summed");
    launch_k1_status(c->stream, c->Lg, c->dClasses, c->in, c->nInst);
    launch_keep_flags(c->stream, c->in, c->nInst, c->dKeep);
    if (c->nInst > 0) launch_exclusive_scan_total(c->stream, c->sched, c->dKeep, c->dKeepOff, c->nInst, c->dScal + 0);
    ctx_mark(c, "update:status, keep scan");
    IBM_CUDA(c, cudaStreamWaitEvent(c->stream, c->updJoin, 0));
    ctx_mark(c, "update:joined");
    if (c->nInst > 0) launch_schedule_chunks(c->stream, c->in, c->nInst, c->sched, c->dScal + 2, c->dScal + 3);
    if (c->v7ok) {
      launch_v7_finish(c->stream, c->in, c->nInst, c->sched, c->v7prep);
      if (comm_ranks(c) > 1) { launch_v7_compact(c->stream, c->nInst, c->sched, c->v7prep, c->dScal + 5); c->nLaunch += 4; }
      c->nLaunch++;
    }
    ctx_mark(c, "update:prepared");
    c->nLaunch += 9;
    IBM_CUDA(c, cudaMemcpyAsync(c->hScal, c->dScal, 6 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    IBM_CUDA(c, cudaMemcpyAsync(c->hScal + 20, c->dScal + 13, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    IBM_CUDA(c, cudaStreamSynchronize(c->stream));
    ctx_mark(c, "update:synced");
    if (c->hScal[20]) return ctx_fail(c, IBM_E_STATE, "staged kernel: a gather stencil left the staged region (internal error)");
    if (!c->hScal[4]) break;
    // shell buffer too small (every rank takes the same path only if it overflows itself: the collective above is
    // re-entered by all ranks or none, because the rerun repeats it) -> grow and redo
    cudaFree(c->dShellCell); cudaFree(c->dShellJK);
    c->shellCap = (int)std::min<long long>(std::max<long long>((long long)c->hScal[1] * 5 / 4 + 4096, (long long)c->shellCap * 2), 0x7fffffff);
    IBM_CUDA(c, cudaMalloc(&c->dShellCell, (size_t)c->shellCap * sizeof(int)));
    IBM_CUDA(c, cudaMalloc(&c->dShellJK, (size_t)c->shellCap * sizeof(int)));
    // (multi-rank: the flag was summed over the ranks above, so every rank is here and repeats the collective)
  }
  if (c->nInst > 0 && c->hScal[0] < c->nInst) {  // deleteInActiveParticles (particleCloud.C:940-950)
    launch_compact_instances(c->stream, c->in, c->in2, c->dKeep, c->dKeepOff, c->nInst);
    std::swap(c->in, c->in2);
    c->nInst = c->hScal[0];
    motion_invalidate_groups(c);
    launch_build_schedule(c->stream, c->L, c->in, c->nInst, c->sched, c->dScal + 2, c->dScal + 3);
    if (c->v7ok) {
      launch_v7_prepare(c->stream, c->L, c->dClasses, c->in, c->nInst, c->sched, c->v7capF, c->v7prep);
      launch_v7_finish(c->stream, c->in, c->nInst, c->sched, c->v7prep);
      if (comm_ranks(c) > 1) { launch_v7_compact(c->stream, c->nInst, c->sched, c->v7prep, c->dScal + 5); c->nLaunch += 4; }
      c->nLaunch += 2;
    }
    c->nLaunch += 6;
    IBM_CUDA(c, cudaMemcpyAsync(c->hScal, c->dScal, 6 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    IBM_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  // several ranks: the staged kernels launch one CTA per instance with local points (1/G of the directory)
  c->v7prep.nLocal = (c->v7ok && comm_ranks(c) > 1) ? c->hScal[5] : -1;
  c->schedN = c->nInst;
  c->totalShell = c->nInst > 0 ? c->hScal[1] : 0;   // extent of the shell buffer in use
  c->totalChunks = c->nInst > 0 ? c->hScal[2] : 0;
  c->updatesPerCall = c->nInst > 0 ? c->hScal[3] : 0;
  // I.ibm (particleCloud::updateI, particleCloud.C:952-969) is not an input of loops 2-5: it is refreshed with the cell
  // lists by definition (App. B #13) but only MATERIALISED when somebody asks for it (ibm_get_indicator) - an 8-byte
  // write per mesh cell and a kernel per update otherwise spent on a field nobody reads
  c->indicatorStale = true;
  ctx_toc(c);
  IBM_CUDA(c, cudaGetLastError());
  return IBM_OK;
}

static int forcing_on_device(ibm_ctx* c, const double* dU, const double* dUold, const double* dS, double dt, double* dFext) {
  const size_t C = (size_t)c->L.nCells;
  double* F = dFext;
  if (!F) {
    int rc = ensure_field(c, &c->dF, 3);
    if (rc) return rc;
    F = c->dF;
    if (!c->fEverZeroed) {
      IBM_CUDA(c, cudaMemsetAsync(F, 0, 3 * C * sizeof(double), c->stream));
      c->fEverZeroed = true;
      c->fLive = false;
    }
    clear_live_F(c);
  } else {
    IBM_CUDA(c, cudaMemsetAsync(F, 0, 3 * C * sizeof(double), c->stream));
  }
  if (c->cfg.coupled) {  // particleCloud.C:1186-1189
    ctx_mark(c, "forcing:begin");
    ctx_tic(c, T_FORCING);
    // 7 = staged kernel (ibm_kernels_v7.cu) where every shape class fits it, 0 = reference-order yardstick, anything else (and
    // clouds the staged kernel cannot take: ellipses, very large bodies) = the direct-gather kernel (ibm_kernels_v2.cu)
    const int fv = c->forcingVariant;
    if (fv >= 7 && c->v7ok)
      launch_forcing_v7(c->stream, c->L, c->dClasses, c->in, c->nInst, c->sched, c->v7prep, c->dShellCell, c->dShellJK, dU, dUold, dS, dt, F,
                        c->v7capF, c->dSlowList, c->dScal + 12, c->dScal + 13);
    else if (fv == 0)
      launch_forcing_ref(c->stream, c->L, c->dClasses, c->in, c->nInst, c->dShellCell, c->dShellJK, dU, dUold, dS, dt, F);
    else
      launch_forcing_v2(c->stream, c->L, c->dClasses, c->in, c->nInst, c->sched, c->totalChunks, c->dShellCell, c->dShellJK, dU, dUold,
                        dS, dt, F);
    c->nLaunch++;
    ctx_toc(c);
    ctx_mark(c, "forcing:end");
    if (!dFext) c->fLive = true;
    c->nForcingCalls++;
    c->nUpdates += c->updatesPerCall;
  }
  IBM_CUDA(c, cudaGetLastError());
  return IBM_OK;
}

// ---- host <-> device transfers of caller-owned arrays ------------------------------------------------------------
// OpenFOAM's field storage is ordinary (pageable) memory unless the caller registered it (ibm_host_register, which the shim
// does for the fields that persist across time steps: U, U.oldTime(), p).  A registered / pinned array is copied directly by
// the DMA engine; a pageable one goes through two library-owned pinned 64 MB chunks, filled by up to sixteen host threads (one per core) while the
// previous chunk is on the bus (cudaMemcpyAsync on pageable memory is a synchronous staged copy at a fraction of the rate).
static bool host_is_pinned(const void* p) {
  cudaPointerAttributes a{};
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}
static size_t chunk_bytes() {
  static size_t v = 0;
  if (!v) {
    const char* e = getenv("IBM_CHUNK_MB");  // developer switch
    const int mb = e ? std::min(1024, std::max(1, std::atoi(e))) : 64;
    v = (size_t)mb << 20;
  }
  return v;
}
#define kChunkBytes chunk_bytes()
static int ensure_chunks(ibm_ctx* c) {
  if (c->hChunk[0]) return IBM_OK;
  for (int q = 0; q < 2; ++q) {
    IBM_CUDA(c, cudaMallocHost(&c->hChunk[q], kChunkBytes));
    IBM_CUDA(c, cudaEventCreateWithFlags(&c->chunkEv[q], cudaEventDisableTiming));
  }
  return IBM_OK;
}
// a few persistent host threads that split one memcpy between them (one pool per process, started on first use)
struct CopyPool {
  int nT = 8;
  std::thread th[32];
  std::mutex mu;
  std::condition_variable cvGo, cvDone;
  long generation = 0;
  int pending = 0;
  bool stop = false;
  char* dst = nullptr; const char* src = nullptr; size_t bytes = 0;
  CopyPool() {
    // measured on the 16-core host of a B200 box (gpurun_out/r2t, 17.2 GB up + 3.2 GB down per step): 4 threads 580 ms per
    // end-to-end step, 8: 528, 16: 453 (49 GB/s on a 55 GB/s link); 16 MB chunks or 256 MB chunks are no better than 64 MB
    if (const char* e = getenv("IBM_COPY_THREADS")) nT = std::min(32, std::max(1, std::atoi(e)));
    else nT = (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
    for (int t = 0; t < nT; ++t) th[t] = std::thread([this, t] { work(t); });
  }
  ~CopyPool() {
    { std::lock_guard<std::mutex> lk(mu); stop = true; }
    cvGo.notify_all();
    for (auto& x : th) if (x.joinable()) x.join();
  }
  void work(int t) {
    long seen = 0;
    for (;;) {
      std::unique_lock<std::mutex> lk(mu);
      cvGo.wait(lk, [&] { return stop || generation != seen; });
      if (stop) return;
      seen = generation;
      char* d = dst; const char* s2 = src; const size_t n = bytes;
      lk.unlock();
      const size_t per = ((n / nT) + 4095) & ~(size_t)4095;
      const size_t o = std::min(n, (size_t)t * per), len = std::min(n - o, per);
      if (len) std::memcpy(d + o, s2 + o, len);
      lk.lock();
      if (--pending == 0) cvDone.notify_all();
    }
  }
  std::mutex callers;  // one copy at a time: contexts of several host threads (in-process ranks) share the pool
  void copy(void* d, const void* s2, size_t n) {
    std::lock_guard<std::mutex> one(callers);
    std::unique_lock<std::mutex> lk(mu);
    dst = (char*)d; src = (const char*)s2; bytes = n; pending = nT; ++generation;
    cvGo.notify_all();
    cvDone.wait(lk, [&] { return pending == 0; });
  }
};
static void parallel_memcpy(void* dst, const void* src, size_t bytes) {
  if (bytes < ((size_t)1 << 20)) { std::memcpy(dst, src, bytes); return; }
  static CopyPool pool;
  pool.copy(dst, src, bytes);
}
static int h2d_bytes(ibm_ctx* c, const void* host, size_t bytes, void* dev) {
  if (host_is_pinned(host)) {
    IBM_CUDA(c, cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, c->stream));
    c->h2dPinnedBytes += (long long)bytes;
    return IBM_OK;
  }
  int rc = ensure_chunks(c);
  if (rc) return rc;
  int k = 0;
  for (size_t o = 0; o < bytes; o += kChunkBytes, ++k) {
    const size_t n = std::min(kChunkBytes, bytes - o);
    const int q = k & 1;
    if (k >= 2) IBM_CUDA(c, cudaEventSynchronize(c->chunkEv[q]));  // the copy that last used this chunk has left it
    parallel_memcpy(c->hChunk[q], (const char*)host + o, n);
    IBM_CUDA(c, cudaMemcpyAsync((char*)dev + o, c->hChunk[q], n, cudaMemcpyHostToDevice, c->stream));
    IBM_CUDA(c, cudaEventRecord(c->chunkEv[q], c->stream));
  }
  IBM_CUDA(c, cudaStreamSynchronize(c->stream));  // the chunks are free again when the call returns
  c->h2dStagedBytes += (long long)bytes;
  return IBM_OK;
}
static int d2h_bytes(ibm_ctx* c, const void* dev, size_t bytes, void* host) {
  if (host_is_pinned(host)) {
    IBM_CUDA(c, cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, c->stream));
    return IBM_OK;
  }
  int rc = ensure_chunks(c);
  if (rc) return rc;
  const size_t nChunks = (bytes + kChunkBytes - 1) / kChunkBytes;
  for (size_t k = 0; k <= nChunks; ++k) {  // chunk k on the bus while chunk k-1 is copied out
    if (k < nChunks) {
      const size_t o = k * kChunkBytes, n = std::min(kChunkBytes, bytes - o);
      IBM_CUDA(c, cudaMemcpyAsync(c->hChunk[k & 1], (const char*)dev + o, n, cudaMemcpyDeviceToHost, c->stream));
      IBM_CUDA(c, cudaEventRecord(c->chunkEv[k & 1], c->stream));
    }
    if (k >= 1) {
      const size_t o = (k - 1) * kChunkBytes, n = std::min(kChunkBytes, bytes - o);
      IBM_CUDA(c, cudaEventSynchronize(c->chunkEv[(k - 1) & 1]));
      parallel_memcpy((char*)host + o, c->hChunk[(k - 1) & 1], n);
    }
  }
  return IBM_OK;
}

static int h2d_field(ibm_ctx* c, const double* hostAoS, int nComp, double* dSoA) {
  const size_t n = (size_t)nComp * c->L.nCells;
  int rc = ctx_ensure_stage(c, n);
  if (rc) return rc;
  ctx_tic(c, T_H2D);
  rc = h2d_bytes(c, hostAoS, n * sizeof(double), c->dStage);
  ctx_toc(c);
  if (rc) return rc;
  c->h2dBytes += (long long)(n * sizeof(double));
  ctx_tic(c, T_LAYOUT);
  launch_aos_to_soa(c->stream, c->L, c->dStage, nComp, dSoA);
  c->nLaunch++;
  ctx_toc(c);
  return IBM_OK;
}

static int d2h_field(ibm_ctx* c, const double* dSoA, int nComp, double* hostAoS) {
  const size_t n = (size_t)nComp * c->L.nCells;
  int rc = ctx_ensure_stage(c, n);
  if (rc) return rc;
  ctx_tic(c, T_LAYOUT);
  launch_soa_to_aos(c->stream, c->L, dSoA, nComp, c->dStage);
  c->nLaunch++;
  ctx_toc(c);
  ctx_tic(c, T_D2H);
  rc = d2h_bytes(c, c->dStage, n * sizeof(double), hostAoS);
  ctx_toc(c);
  if (rc) return rc;
  c->d2hBytes += (long long)(n * sizeof(double));
  return IBM_OK;
}

static int stress_on_device(ibm_ctx* c, const double* dP, const double* dTau, double* FT_host) {
  double pRefOverRho = c->cfg.pRef / c->cfg.rhoRef;  // particleCloud.C:1213-1215
  if (!c->dTcorr) IBM_CUDA(c, cudaMalloc(&c->dTcorr, (size_t)3 * c->cap * sizeof(double)));
  ctx_tic(c, T_STRESS);
  const int sv = c->stressVariant;
  if (sv >= 7 && c->v7ok) {
    const bool clr = c->clearFInStress && c->fLive && c->dF && c->nInst > 0;
    launch_stress_v7(c->stream, c->L, c->dClasses, c->in, c->nInst, c->sched, c->v7prep, dP, dTau, pRefOverRho, c->cfg.rhoRef, c->dTcorr, c->v7capS,
                     c->dSlowList, c->dScal + 12, c->dScal + 13, clr ? c->dF : nullptr, c->dShellCell);
    if (clr) c->fLive = false;
  }
  else if (sv == 0)
    launch_stress_ref(c->stream, c->L, c->dClasses, c->in, c->nInst, dP, dTau, pRefOverRho, c->cfg.rhoRef, c->dTcorr);
  else
    launch_stress_v2(c->stream, c->L, c->dClasses, c->in, c->nInst, c->sched, dP, dTau, pRefOverRho, c->cfg.rhoRef, c->dTcorr);
  c->nLaunch++;
  ctx_toc(c);
  IBM_CUDA(c, cudaGetLastError());
  if (FT_host) {
    const int n = c->nInst;
    int rc = ctx_ensure_pinned(c, (size_t)6 * n + 16);
    if (rc) return rc;
    ctx_tic(c, T_D2H);
    for (int q = 0; q < 3; ++q) {
      IBM_CUDA(c, cudaMemcpyAsync(c->hPinned + (size_t)q * n, c->in.F[q], n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
      IBM_CUDA(c, cudaMemcpyAsync(c->hPinned + (size_t)(3 + q) * n, c->in.T[q], n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    }
    ctx_toc(c);
    c->d2hBytes += (long long)(6 * n * sizeof(double));
    IBM_CUDA(c, cudaStreamSynchronize(c->stream));
    for (int i = 0; i < n; ++i)
      for (int q = 0; q < 6; ++q) FT_host[6 * (size_t)i + q] = c->hPinned[(size_t)q * n + i];
  }
  return IBM_OK;
}

}  // namespace ibm

// =================================================================================================
// extern "C"
// =================================================================================================
extern "C" {

int ibm_abi_version(void) { return IBM_ABI_VERSION; }

const char* ibm_last_error(const ibm_ctx* ctx) { return ctx ? ctx->err.c_str() : g_createErr.c_str(); }

int ibm_create(const ibm_mesh_desc* mesh, const ibm_cloud_cfg* cfg, int device, ibm_ctx** out) {
  if (!mesh || !cfg || !out) return ctx_fail(nullptr, IBM_E_ARG, "ibm_create: null argument");
  *out = nullptr;
  if (mesh->nx <= 0 || mesh->ny <= 0 || mesh->nz <= 0) return ctx_fail(nullptr, IBM_E_ARG, "ibm_create: empty lattice");
  if ((long long)mesh->nx * mesh->ny * mesh->nz > 0x7fffffffLL / 3 * 3 && (long long)mesh->nx * mesh->ny * mesh->nz > 0x7fffffffLL)
    return ctx_fail(nullptr, IBM_E_ARG, "ibm_create: lattice exceeds 32-bit labels");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return ctx_fail(nullptr, IBM_E_CUDA, "ibm_create: no CUDA device (this path has no CPU fallback)");
  if (device < 0 || device >= ndev) return ctx_fail(nullptr, IBM_E_ARG, "ibm_create: bad device ordinal");
  if (!ctx_cuda_ok(nullptr, cudaSetDevice(device), "cudaSetDevice")) return IBM_E_CUDA;
  ibm_ctx* c = new ibm_ctx;
  c->device = device;
  c->cfg = *cfg;
  if (const char* e = getenv("IBM_FUSED_LOOP5")) c->fusedLoop5 = std::atoi(e) != 0;  // developer switches (profiling)
  if (const char* e = getenv("IBM_FUSED_ACTIVE_LOOP5")) c->fusedActive = std::atoi(e) != 0;
  if (const char* e = getenv("IBM_GRAPH_LOOP5")) c->graphLoop5 = std::atoi(e) != 0;
  if (const char* e = getenv("IBM_CHUNK256")) c->chunk256 = std::atoi(e) != 0;
  if (const char* e = getenv("IBM_ACTIVE_LOOP5")) c->activeLoop5 = std::atoi(e) != 0;
  if (const char* e = getenv("IBM_CLUSTER_LOOP5")) c->clusterLoop5 = std::atoi(e) != 0;
  if (const char* e = getenv("IBM_OWNED_LOOP5")) c->ownedLoop5 = std::atoi(e) != 0;
  if (const char* e = getenv("IBM_DEBUG_TIMING")) c->dbgTiming = std::atoi(e) != 0;
  if (const char* e = getenv("IBM_MARK_VARIANT")) c->markVariant = std::atoi(e) == 1 ? 1 : 0;
  if (const char* e = getenv("IBM_KERNEL_VARIANT")) {  // "f,s": forcing / stress kernel generation (like ibm_set_kernel_variant)
    int f = 0, s2 = 0;
    if (std::sscanf(e, "%d,%d", &f, &s2) == 2) { c->forcingVariant = f; c->stressVariant = s2; }
  }
  if (!ctx_cuda_ok(nullptr, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking), "cudaStreamCreate")) { delete c; return IBM_E_CUDA; }
  LatticeDev& L = c->L;
  L.nx = mesh->nx; L.ny = mesh->ny; L.nz = mesh->nz;
  L.nCells = mesh->nx * mesh->ny * mesh->nz;
  L.ox = mesh->origin[0]; L.oy = mesh->origin[1]; L.oz = mesh->origin[2];
  L.dx = mesh->delta[0]; L.dy = mesh->delta[1]; L.dz = mesh->delta[2];
  L.latToCell = nullptr; L.cellToLat = nullptr;
  L.kz0 = 0;
  L.ownK0 = 0; L.ownK1 = L.nz;
  c->zMin = L.oz;
  c->zMax = L.oz + L.nz * L.dz;
  // caller cell numbering
  const size_t n = (size_t)L.nCells;
  bool haveMap = false;
  if (mesh->latticeToCell) {
    c->hLatToCell.assign(mesh->latticeToCell, mesh->latticeToCell + n);
    haveMap = true;
  } else if (mesh->nBlocks > 1 || (mesh->nBlocks == 1 && mesh->blocks)) {
    c->hLatToCell.assign(n, -1);
    int next = 0;
    for (int b = 0; b < mesh->nBlocks; ++b) {
      const int* q = mesh->blocks + 6 * b;
      for (int K = q[4]; K < q[5]; ++K)
        for (int J = q[2]; J < q[3]; ++J)
          for (int I = q[0]; I < q[1]; ++I) {
            if (I < 0 || I >= L.nx || J < 0 || J >= L.ny || K < 0 || K >= L.nz) { delete c; return ctx_fail(nullptr, IBM_E_ARG, "ibm_create: block outside lattice"); }
            c->hLatToCell[I + (size_t)L.nx * (J + (size_t)L.ny * K)] = next++;
          }
    }
    if ((size_t)next != n || std::find(c->hLatToCell.begin(), c->hLatToCell.end(), -1) != c->hLatToCell.end()) {
      delete c;
      return ctx_fail(nullptr, IBM_E_ARG, "ibm_create: blocks do not tile the lattice");
    }
    haveMap = mesh->nBlocks > 1;
    if (!haveMap) c->hLatToCell.clear();
  }
  if (haveMap) {
    std::vector<int> inv(n, -1);
    for (size_t l = 0; l < n; ++l) {
      int id = c->hLatToCell[l];
      if (id < 0 || (size_t)id >= n || inv[id] != -1) { delete c; return ctx_fail(nullptr, IBM_E_ARG, "ibm_create: latticeToCell is not a permutation"); }
      inv[id] = (int)l;
    }
    if (!ctx_cuda_ok(nullptr, cudaMalloc(&c->dLatToCell, n * sizeof(int)), "cudaMalloc") ||
        !ctx_cuda_ok(nullptr, cudaMalloc(&c->dCellToLat, n * sizeof(int)), "cudaMalloc")) { delete c; return IBM_E_CUDA; }
    cudaMemcpy(c->dLatToCell, c->hLatToCell.data(), n * sizeof(int), cudaMemcpyHostToDevice);
    cudaMemcpy(c->dCellToLat, inv.data(), n * sizeof(int), cudaMemcpyHostToDevice);
    L.latToCell = c->dLatToCell; L.cellToLat = c->dCellToLat;
  }
  // patches (particleCloud.C:1050-1068)
  bool emptyDir[3] = {false, false, false};
  for (int p = 0; p < mesh->nPatches; ++p) {
    const ibm_patch_desc& pd = mesh->patches[p];
    PatchHost ph;
    ph.type = pd.type; ph.side = pd.side; ph.neighb = pd.neighbPatch;
    if (pd.side < 0 || pd.side > 5) { delete c; return ctx_fail(nullptr, IBM_E_ARG, "ibm_create: bad patch side"); }
    const int axis = pd.side / 2;
    const bool hi = pd.side & 1;
    if (pd.type == IBM_EMPTY) emptyDir[axis] = true;
    if (pd.Cf && pd.Sf && pd.magSf) {
      ph.Cf.assign(pd.Cf, pd.Cf + 3 * (size_t)pd.nFaces);
      ph.Sf.assign(pd.Sf, pd.Sf + 3 * (size_t)pd.nFaces);
      ph.magSf.assign(pd.magSf, pd.magSf + pd.nFaces);
    } else {
      // mini-mesh patch faces: transverse lattice order, fastest index first
      const int nA = (axis == 0) ? L.ny : L.nx, nB = (axis == 2) ? L.ny : L.nz;
      const double d3[3] = {L.dx, L.dy, L.dz}, o3[3] = {L.ox, L.oy, L.oz};
      const int n3[3] = {L.nx, L.ny, L.nz};
      for (int b = 0; b < nB; ++b)
        for (int a = 0; a < nA; ++a) {
          int ijk[3];
          if (axis == 0) { ijk[0] = hi ? L.nx - 1 : 0; ijk[1] = a; ijk[2] = b; }
          else if (axis == 1) { ijk[0] = a; ijk[1] = hi ? L.ny - 1 : 0; ijk[2] = b; }
          else { ijk[0] = a; ijk[1] = b; ijk[2] = hi ? L.nz - 1 : 0; }
          double cf[3], sf[3] = {0, 0, 0};
          for (int q = 0; q < 3; ++q) cf[q] = o3[q] + (ijk[q] + 0.5) * d3[q];
          cf[axis] = hi ? o3[axis] + n3[axis] * d3[axis] : o3[axis];
          double area = (axis == 0) ? d3[1] * d3[2] : (axis == 1 ? d3[0] * d3[2] : d3[0] * d3[1]);
          sf[axis] = (hi ? 1.0 : -1.0) * area;
          for (int q = 0; q < 3; ++q) { ph.Cf.push_back(cf[q]); ph.Sf.push_back(sf[q]); }
          ph.magSf.push_back(std::sqrt(sf[0] * sf[0] + sf[1] * sf[1] + sf[2] * sf[2]));
        }
    }
    if (ph.nFaces() > 0) {
      size_t nf = (size_t)ph.nFaces();
      if (!ctx_cuda_ok(nullptr, cudaMalloc(&ph.dCf, 3 * nf * sizeof(double)), "cudaMalloc") ||
          !ctx_cuda_ok(nullptr, cudaMalloc(&ph.dSf, 3 * nf * sizeof(double)), "cudaMalloc") ||
          !ctx_cuda_ok(nullptr, cudaMalloc(&ph.dMagSf, nf * sizeof(double)), "cudaMalloc")) { delete c; return IBM_E_CUDA; }
      cudaMemcpy(ph.dCf, ph.Cf.data(), 3 * nf * sizeof(double), cudaMemcpyHostToDevice);
      cudaMemcpy(ph.dSf, ph.Sf.data(), 3 * nf * sizeof(double), cudaMemcpyHostToDevice);
      cudaMemcpy(ph.dMagSf, ph.magSf.data(), nf * sizeof(double), cudaMemcpyHostToDevice);
    }
    if ((pd.type == IBM_CYCLIC || pd.type == IBM_PROCESSOR) && ph.nFaces() != 0) c->cyclics.push_back(p);
    else if (pd.type == IBM_WALL) c->walls.push_back(p);
    c->patches.push_back(std::move(ph));
  }
  c->nGeometricD = 3 - (int)emptyDir[0] - (int)emptyDir[1] - (int)emptyDir[2];
  c->Lg = c->L;
  if (!ctx_cuda_ok(nullptr, cudaMalloc(&c->dScal, 64 * sizeof(int)), "cudaMalloc") ||
      !ctx_cuda_ok(nullptr, cudaMallocHost(&c->hScal, 64 * sizeof(int)), "cudaMallocHost")) { delete c; return IBM_E_CUDA; }
  cudaMemset(c->dScal, 0, 64 * sizeof(int));
  *out = c;
  return IBM_OK;
}

int ibm_set_particles(ibm_ctx* c, int n, const ibm_particle_init* init) {
  if (!c || n < 0 || (n > 0 && !init)) return ctx_fail(c, IBM_E_ARG, "ibm_set_particles: bad arguments");
  cudaSetDevice(c->device);
  if (c->cap != 0) return ctx_fail(c, IBM_E_STATE, "ibm_set_particles: particles already set");
  c->nParticles = n;
  c->orig.assign(init, init + n);
  c->origCls.assign(n, -1);
  c->maxCopies.assign(n, -1);
  const double l = std::fabs(c->zMax - c->zMin);  // cylinder.C:73, ellipse.C:90
  for (int i = 0; i < n; ++i) {
    const ibm_particle_init& p = init[i];
    // sphere.C:79-84, cylinder.C:64-70, ellipse.C:80-86: FatalError in the reference
    if (p.kind == IBM_SPHERE && c->nGeometricD != 3) return ctx_fail(c, IBM_E_ARG, "Sphere particle shape only valid for 3D meshes");
    if ((p.kind == IBM_CYLINDER || p.kind == IBM_ELLIPSE) && c->nGeometricD != 2)
      return ctx_fail(c, IBM_E_ARG, "cylinder/ellipse particle shape with 1 point in axis only valid for 2D meshes");
    if (p.kind < 0 || p.kind > 2) return ctx_fail(c, IBM_E_ARG, "Unknown particleShape type");  // particleShapeNew.C:44-49
    if (p.nTheta < 2 || (p.kind == IBM_SPHERE && p.nTheta < 4) || p.nTheta > 65535) return ctx_fail(c, IBM_E_ARG, "nTheta out of range");
    int found = -1;
    for (size_t k = 0; k < c->classes.size(); ++k) {
      const ShapeClassDev& s = c->classes[k].dev;
      if (s.kind == p.kind && s.nTheta == p.nTheta && s.delta == p.delta &&
          ((p.kind == IBM_ELLIPSE) ? (s.a == p.a && s.b == p.b) : (s.d == p.d))) { found = (int)k; break; }
    }
    if (found < 0) {
      c->classes.emplace_back();
      build_class(c->classes.back(), p.kind, p.d, p.a, p.b, p.nTheta, p.delta, l);
      found = (int)c->classes.size() - 1;
    }
    c->origCls[i] = found;
  }
  int rc = upload_classes(c);
  if (rc) return rc;
  {
    std::vector<ShapeClassDev> hc;
    for (auto& h : c->classes) hc.push_back(h.dev);
    c->v7ok = !hc.empty() && v7_plan(c->L, hc.data(), (int)hc.size(), c->nGeometricD, c->v7capF, c->v7capS);
  }
  c->cap = c->cfg.maxInstances > 0 ? std::max(c->cfg.maxInstances, n) : 2 * n + 64;
  IBM_CUDA(c, cudaMalloc(&c->dSlowList, (size_t)c->cap * sizeof(int)));
  if (c->v7ok && v7_prep_alloc(c->v7prep, c->cap, std::max(c->v7capF.rowsCap, c->v7capS.rowsCap))) return ctx_fail(c, IBM_E_CUDA, "v7_prep_alloc failed");
  if ((rc = alloc_instances(c, c->in, c->cap))) return rc;
  if ((rc = alloc_instances(c, c->in2, c->cap))) return rc;
  if (schedule_alloc(c->sched, c->cap)) return ctx_fail(c, IBM_E_CUDA, "schedule_alloc failed");
  IBM_CUDA(c, cudaMalloc(&c->dKeep, ((size_t)c->cap + 1) * sizeof(int)));
  IBM_CUDA(c, cudaMalloc(&c->dKeepOff, ((size_t)c->cap + 1) * sizeof(int)));
  // upload originals: particleIBM(mesh, dict, index)  particleIBM.C:269-315
  std::vector<int> iv(n);
  std::vector<double> dv(n);
  auto upI = [&](int* dst, auto f) { for (int i = 0; i < n; ++i) iv[i] = f(i); cudaMemcpy(dst, iv.data(), n * sizeof(int), cudaMemcpyHostToDevice); };
  auto upD = [&](double* dst, auto f) { for (int i = 0; i < n; ++i) dv[i] = f(i); cudaMemcpy(dst, dv.data(), n * sizeof(double), cudaMemcpyHostToDevice); };
  if (n > 0) {
    upI(c->in.index, [&](int i) { return i; });
    upI(c->in.copy, [&](int) { return 0; });
    upI(c->in.active, [&](int) { return 1; });
    upI(c->in.onMesh, [&](int) { return IBM_ON_MESH; });
    upI(c->in.cls, [&](int i) { return c->origCls[i]; });
    for (int q = 0; q < 3; ++q) {
      upD(c->in.c[q], [&](int i) { return init[i].position[q]; });
      upD(c->in.th[q], [&](int i) { return init[i].theta[q]; });
      upD(c->in.v[q], [&](int i) { return init[i].v[q]; });
      upD(c->in.om[q], [&](int i) { return init[i].omega[q]; });
    }
    upD(c->in.rho, [&](int i) { return init[i].rho; });
  }
  c->nInst = n;
  rc = ctx_update(c);
  if (rc) return rc;
  // maxCopies_[index] = 0 for every surviving particle (particleCloud.C:1127-1131)
  if (c->nInst > 0) {
    std::vector<int> idx(c->nInst);
    cudaMemcpy(idx.data(), c->in.index, c->nInst * sizeof(int), cudaMemcpyDeviceToHost);
    for (int i : idx) c->maxCopies[i] = 0;
  }
  return IBM_OK;
}

int ibm_destroy(ibm_ctx* c) {
  if (!c) return IBM_OK;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  ctx_marks_report(c);
  motion_destroy(c);
  comm_destroy(c);
  schedule_free(c->sched);
  v7_prep_free(c->v7prep);
  if (c->updSide) { cudaStreamSynchronize(c->updSide); cudaStreamDestroy(c->updSide); cudaEventDestroy(c->updFork); cudaEventDestroy(c->updJoin); }
  for (auto& p : c->events) { cudaEventDestroy(p.a); if (p.b) cudaEventDestroy(p.b); }
  for (auto e : c->eventPool) cudaEventDestroy(e);
  for (auto p : c->instAllocs) cudaFree(p);
  for (auto p : c->classPools) cudaFree(p);
  for (auto& ph : c->patches) { cudaFree(ph.dCf); cudaFree(ph.dSf); cudaFree(ph.dMagSf); }
  void* frees[] = {c->dLatToCell, c->dCellToLat, c->dClasses, c->dShellCell, c->dShellJK, c->dScal, c->dKeep, c->dKeepOff,
                   c->dU, c->dUold, c->dS, c->dF, c->dP, c->dTau, c->dI, c->dStage, c->dTcorr, c->dSlowList,
                   c->dNearFlag, c->dNearOff, c->dNearList};
  for (void* p : frees) if (p) cudaFree(p);
  if (c->hScal) cudaFreeHost(c->hScal);
  if (c->hPinned) cudaFreeHost(c->hPinned);
  for (int q = 0; q < 2; ++q) { if (c->hChunk[q]) cudaFreeHost(c->hChunk[q]); if (c->chunkEv[q]) cudaEventDestroy(c->chunkEv[q]); }
  cudaStreamDestroy(c->stream);
  delete c;
  return IBM_OK;
}

int ibm_lattice_from_arrays(const double* C, int nCells, int nGeometricD, int nPatches, const double* const* patchCf, const int* patchNFaces,
                            int* n, double* origin, double* delta, int* patchSide, int* latticeToCell, int* identity) {
  if (!C || nCells <= 0 || !n || !origin || !delta || (nPatches > 0 && (!patchCf || !patchNFaces || !patchSide)))
    return ctx_fail(nullptr, IBM_E_ARG, "ibm_lattice_from_arrays: bad arguments");
  double lo[3], hi[3], step[3];
  for (int d = 0; d < 3; ++d) { lo[d] = 1e300; hi[d] = -1e300; }
  for (int c = 0; c < nCells; ++c)
    for (int d = 0; d < 3; ++d) { lo[d] = std::min(lo[d], C[3 * (size_t)c + d]); hi[d] = std::max(hi[d], C[3 * (size_t)c + d]); }
  const double span = std::max(std::max(hi[0] - lo[0], hi[1] - lo[1]), std::max(hi[2] - lo[2], 1e-300));
  long long nn[3];
  for (int d = 0; d < 3; ++d) {
    // spacing = the smallest centre coordinate above the lowest one
    double s = 1e300;
    const double tol = 1e-9 * span;
    for (int c = 0; c < nCells; ++c) { const double x = C[3 * (size_t)c + d] - lo[d]; if (x > tol && x < s) s = x; }
    if (s > 1e299) { step[d] = 0.0; nn[d] = 1; continue; }   // one layer of cells along d
    step[d] = s;
    nn[d] = (long long)std::floor((hi[d] - lo[d]) / s + 0.5) + 1;
  }
  if (nn[0] * nn[1] * nn[2] != (long long)nCells)
    return ctx_fail(nullptr, IBM_E_UNSUPPORTED, "ibm_lattice_from_arrays: the cell centres are not a uniform lattice (graded or unstructured mesh)");
  // a single layer has no spacing of its own: the caller's empty direction (2-D cases); take the spacing of x
  for (int d = 0; d < 3; ++d)
    if (nn[d] == 1) step[d] = (nGeometricD == 2 && step[0] > 0) ? step[0] : (step[0] > 0 ? step[0] : 1.0);
  std::vector<char> seen((size_t)nCells, 0);
  bool ident = true;
  for (int c = 0; c < nCells; ++c) {
    long long idx[3];
    for (int d = 0; d < 3; ++d) {
      const double t = (C[3 * (size_t)c + d] - lo[d]) / step[d];
      idx[d] = (long long)std::floor(t + 0.5);
      if (idx[d] < 0 || idx[d] >= nn[d] || std::fabs(t - (double)idx[d]) > 1e-6)
        return ctx_fail(nullptr, IBM_E_UNSUPPORTED, "ibm_lattice_from_arrays: the cell centres are not a uniform lattice (graded or unstructured mesh)");
    }
    const size_t lat = (size_t)(idx[0] + nn[0] * (idx[1] + nn[1] * idx[2]));
    if (seen[lat]) return ctx_fail(nullptr, IBM_E_UNSUPPORTED, "ibm_lattice_from_arrays: two cells share a lattice position");
    seen[lat] = 1;
    if (latticeToCell) latticeToCell[lat] = c;
    ident = ident && lat == (size_t)c;
  }
  for (int d = 0; d < 3; ++d) { n[d] = (int)nn[d]; delta[d] = step[d]; origin[d] = lo[d] - 0.5 * step[d]; }
  // the thickness of a single layer is not in the cell centres: it is the distance between the two patches that bound it
  for (int d = 0; d < 3; ++d) {
    if (nn[d] != 1) continue;
    double below = -1e300, above = 1e300;
    for (int p = 0; p < nPatches; ++p) {
      const int nf = patchNFaces[p];
      if (nf <= 0) continue;
      const double v = patchCf[p][d];
      bool flat = true;
      for (int f = 1; f < nf && flat; ++f) flat = std::fabs(patchCf[p][3 * (size_t)f + d] - v) <= 1e-9 * span;
      if (!flat) continue;
      if (v < lo[d]) below = std::max(below, v);
      if (v > lo[d]) above = std::min(above, v);
    }
    if (below > -1e299 && above < 1e299) { origin[d] = below; delta[d] = above - below; }
  }
  if (identity) *identity = ident ? 1 : 0;
  for (int p = 0; p < nPatches; ++p) {
    patchSide[p] = -1;
    const int nf = patchNFaces[p];
    if (nf <= 0) { patchSide[p] = IBM_XMIN; continue; }  // (an empty patch hits nothing: any side)
    for (int side = 0; side < 6 && patchSide[p] < 0; ++side) {
      const int d = side / 2;
      const double plane = (side & 1) ? origin[d] + n[d] * delta[d] : origin[d];
      bool all = true;
      for (int f = 0; f < nf && all; ++f) all = std::fabs(patchCf[p][3 * (size_t)f + d] - plane) <= 1e-6 * delta[d];
      if (all) patchSide[p] = side;
    }
    if (patchSide[p] < 0) return ctx_fail(nullptr, IBM_E_UNSUPPORTED, "ibm_lattice_from_arrays: a boundary patch does not lie on one side of the lattice");
  }
  return IBM_OK;
}

int ibm_set_slab_halo(ibm_ctx* c, int k0, int k1, int halo, int* localK0, int* localK1) {
  if (!c) return IBM_E_ARG;
  if (c->cap != 0) return ctx_fail(c, IBM_E_STATE, "ibm_set_slab: call before ibm_set_particles");
  if (c->Lg.latToCell) return ctx_fail(c, IBM_E_UNSUPPORTED, "ibm_set_slab: slab decomposition needs a single-block lattice");
  if (k0 < 0 || k1 > c->Lg.nz || k0 >= k1 || halo < 0) return ctx_fail(c, IBM_E_ARG, "ibm_set_slab: bad plane range");
  const int lo = std::max(0, k0 - halo), hi = std::min(c->Lg.nz, k1 + halo);
  c->L = c->Lg;
  c->L.kz0 = lo;
  c->L.nz = hi - lo;
  c->L.ownK0 = k0 - lo;
  c->L.ownK1 = k1 - lo;
  c->L.nCells = c->L.nx * c->L.ny * c->L.nz;
  if (localK0) *localK0 = lo;
  if (localK1) *localK1 = hi;
  return IBM_OK;
}

int ibm_set_slab(ibm_ctx* c, int k0, int k1) { return ibm_set_slab_halo(c, k0, k1, 0, nullptr, nullptr); }

int ibm_update(ibm_ctx* c) {
  if (!c) return IBM_E_ARG;
  cudaSetDevice(c->device);
  return ctx_update(c);
}

int ibm_forcing(ibm_ctx* c, const double* U, const double* Uold, const double* S, double dt, double* F) {
  if (!c || !U || !Uold || !S || !F) return ctx_fail(c, IBM_E_ARG, "ibm_forcing: null argument");
  cudaSetDevice(c->device);
  const size_t C = (size_t)c->L.nCells;
  if (!c->cfg.coupled) {  // particleCloud.C:1186-1189: a zero field
    std::memset(F, 0, 3 * C * sizeof(double));
    return IBM_OK;
  }
  int rc;
  if ((rc = ensure_field(c, &c->dU, 3)) || (rc = ensure_field(c, &c->dUold, 3)) || (rc = ensure_field(c, &c->dS, 3))) return rc;
  if ((rc = h2d_field(c, U, 3, c->dU)) || (rc = h2d_field(c, Uold, 3, c->dUold)) || (rc = h2d_field(c, S, 3, c->dS))) return rc;
  if ((rc = forcing_on_device(c, c->dU, c->dUold, c->dS, dt, nullptr))) return rc;
  if ((rc = d2h_field(c, c->dF, 3, F))) return rc;
  IBM_CUDA(c, cudaStreamSynchronize(c->stream));
  return IBM_OK;
}

int ibm_forcing_device(ibm_ctx* c, const double* dU, const double* dUold, const double* dS, double dt, double* dF) {
  if (!c || !dU || !dUold || !dS) return ctx_fail(c, IBM_E_ARG, "ibm_forcing_device: null argument");
  cudaSetDevice(c->device);
  return forcing_on_device(c, dU, dUold, dS, dt, dF);
}

int ibm_integrate_stress(ibm_ctx* c, const double* p, const double* tau6, double* FT) {
  if (!c || !p || !tau6) return ctx_fail(c, IBM_E_ARG, "ibm_integrate_stress: null argument");
  cudaSetDevice(c->device);
  int rc;
  if ((rc = ensure_field(c, &c->dP, 1)) || (rc = ensure_field(c, &c->dTau, 6))) return rc;
  if ((rc = h2d_field(c, p, 1, c->dP)) || (rc = h2d_field(c, tau6, 6, c->dTau))) return rc;
  if ((rc = stress_on_device(c, c->dP, c->dTau, FT))) return rc;
  IBM_CUDA(c, cudaStreamSynchronize(c->stream));
  return IBM_OK;
}

int ibm_integrate_stress_device(ibm_ctx* c, const double* dP, const double* dTau6, double* FT_host) {
  if (!c || !dP || !dTau6) return ctx_fail(c, IBM_E_ARG, "ibm_integrate_stress_device: null argument");
  cudaSetDevice(c->device);
  return stress_on_device(c, dP, dTau6, FT_host);
}

int ibm_combine_forces(ibm_ctx* c) {
  if (!c) return IBM_E_ARG;
  cudaSetDevice(c->device);
  return combine_forces(c);
}

int ibm_advance(ibm_ctx* c, double dt, ibm_solve_stats* stats) {
  if (!c) return IBM_E_ARG;
  cudaSetDevice(c->device);
  return motion_advance(c, dt, stats);
}

int ibm_substep(ibm_ctx* c, double localDt) {
  if (!c) return IBM_E_ARG;
  cudaSetDevice(c->device);
  return motion_substep(c, localDt);
}

// the sparse re-zeroing of the forcing field (its consumer is done with it once solve() is called) beside the stress integral
static void clear_F_on_side(ibm_ctx* c) {
  cudaStream_t side = motion_side_stream(c);
  if (!side || !c->fLive || !c->dF) return;
  if (c->stressVariant >= 7 && c->v7ok) return;  // the staged stress kernel does it on the way (clearFInStress)
  launch_clear_shell_cells(side, c->dShellCell, c->totalShell, (size_t)c->L.nCells, c->dF);
  c->nLaunch++;
  c->fLive = false;
  // (the side stream's work is joined by motion_advance before the shell lists are rebuilt)
}

int ibm_solve(ibm_ctx* c, const double* p, const double* tau6, double dt, ibm_solve_stats* stats) {
  if (!c) return IBM_E_ARG;
  cudaSetDevice(c->device);
  int rc = motion_prepare_async(c, dt);
  if (rc) return rc;
  clear_F_on_side(c);
  c->clearFInStress = true;
  rc = ibm_integrate_stress(c, p, tau6, nullptr);
  c->clearFInStress = false;
  if (rc) return rc;
  if ((rc = combine_forces(c))) return rc;
  return motion_advance(c, dt, stats);
}

int ibm_solve_device(ibm_ctx* c, const double* dP, const double* dTau6, double dt, ibm_solve_stats* stats) {
  if (!c) return IBM_E_ARG;
  cudaSetDevice(c->device);
  ctx_mark(c, "solve:begin");
  int rc = motion_prepare_async(c, dt);
  if (rc) return rc;
  clear_F_on_side(c);
  ctx_mark(c, "solve:prepared");
  c->clearFInStress = true;
  rc = ibm_integrate_stress_device(c, dP, dTau6, nullptr);
  c->clearFInStress = false;
  if (rc) return rc;
  ctx_mark(c, "solve:stress done");
  if ((rc = combine_forces(c))) return rc;
  ctx_mark(c, "solve:combined");
  rc = motion_advance(c, dt, stats);
  ctx_mark(c, "solve:end");
  return rc;
}

int ibm_n_instances(ibm_ctx* c, int* n) {
  if (!c || !n) return IBM_E_ARG;
  *n = c->nInst;
  return IBM_OK;
}

// the accessor records assembled on the device (one thread per instance), so that the host gets them with ONE copy
__global__ void k_pack_particle_state(InstanceSoA in, const ShapeClassDev* __restrict__ classes, int n, ibm_particle_state* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  ibm_particle_state s;
  s.index = in.index[i]; s.copy = in.copy[i]; s.active = in.active[i]; s.onMesh = in.onMesh[i];
  for (int q = 0; q < 3; ++q) {
    s.center[q] = in.c[q][i]; s.theta[q] = in.th[q][i]; s.v[q] = in.v[q][i]; s.omega[q] = in.om[q][i];
    s.F[q] = in.F[q][i]; s.T[q] = in.T[q][i];
  }
  s.rho = in.rho[i]; s.age = in.age[i];
  s.nShell = in.nShell[i]; s.N = classes[in.cls[i]].N;
  out[i] = s;
}

int ibm_get_particles(ibm_ctx* c, ibm_particle_state* out) {
  if (!c || (!out && c->nInst > 0)) return ctx_fail(c, IBM_E_ARG, "ibm_get_particles: null argument");
  cudaSetDevice(c->device);
  const int n = c->nInst;
  if (n == 0) return IBM_OK;
  const size_t bytes = (size_t)n * sizeof(ibm_particle_state);
  int rc = ctx_ensure_stage(c, (bytes + sizeof(double) - 1) / sizeof(double));
  if (rc) return rc;
  ibm_particle_state* d = (ibm_particle_state*)c->dStage;
  k_pack_particle_state<<<(n + 127) / 128, 128, 0, c->stream>>>(c->in, c->dClasses, n, d);
  c->nLaunch++;
  IBM_CUDA(c, cudaMemcpyAsync(out, d, bytes, cudaMemcpyDeviceToHost, c->stream));
  IBM_CUDA(c, cudaStreamSynchronize(c->stream));
  IBM_CUDA(c, cudaGetLastError());
  return IBM_OK;
}

int ibm_restore_particles(ibm_ctx* c, int n, const ibm_particle_state* st) {
  if (!c || n < 0 || (n > 0 && !st)) return ctx_fail(c, IBM_E_ARG, "ibm_restore_particles: bad arguments");
  if (c->cap == 0) return ctx_fail(c, IBM_E_STATE, "ibm_restore_particles: call ibm_set_particles first (origParticles_)");
  if (n > c->cap) return ctx_fail(c, IBM_E_CAPACITY, "ibm_restore_particles: more instances than ibm_cloud_cfg.maxInstances");
  cudaSetDevice(c->device);
  IBM_CUDA(c, cudaStreamSynchronize(c->stream));
  for (int i = 0; i < n; ++i)
    if (st[i].index < 0 || st[i].index >= c->nParticles) return ctx_fail(c, IBM_E_ARG, "ibm_restore_particles: index outside IBMProperties");
  clear_live_F(c);  // (the shell lists of the instances that go away still say where F is non-zero)
  std::vector<int> iv(std::max(n, 1));
  std::vector<double> dv(std::max(n, 1));
  auto upI = [&](int* dst, auto f) { for (int i = 0; i < n; ++i) iv[i] = f(i); cudaMemcpy(dst, iv.data(), n * sizeof(int), cudaMemcpyHostToDevice); };
  auto upD = [&](double* dst, auto f) { for (int i = 0; i < n; ++i) dv[i] = f(i); cudaMemcpy(dst, dv.data(), n * sizeof(double), cudaMemcpyHostToDevice); };
  if (n > 0) {
    upI(c->in.index, [&](int i) { return st[i].index; });
    upI(c->in.copy, [&](int i) { return st[i].copy; });
    upI(c->in.active, [&](int i) { return st[i].active; });
    upI(c->in.onMesh, [&](int) { return IBM_ON_MESH; });
    upI(c->in.cls, [&](int i) { return c->origCls[st[i].index]; });
    upI(c->in.nShell, [&](int) { return 0; });
    upI(c->in.shellOff, [&](int) { return 0; });
    for (int q = 0; q < 3; ++q) {
      upD(c->in.c[q], [&](int i) { return st[i].center[q]; });
      upD(c->in.th[q], [&](int i) { return st[i].theta[q]; });
      upD(c->in.v[q], [&](int i) { return st[i].v[q]; });
      upD(c->in.om[q], [&](int i) { return st[i].omega[q]; });
      upD(c->in.F[q], [&](int) { return 0.0; });
      upD(c->in.T[q], [&](int) { return 0.0; });
    }
    upD(c->in.rho, [&](int i) { return st[i].rho; });
    upD(c->in.age, [&](int i) { return st[i].age; });
  }
  c->nInst = n;
  c->totalShell = 0;
  motion_invalidate_groups(c);
  int rc = ctx_update(c);  // update() + deleteInActiveParticles (particleCloud.C:1125)
  if (rc) return rc;
  // AS CODED: particleIBM::readFields builds every shape with particleShape::New(mesh, particle.<index>, center) - i.e. with the
  // DICTIONARY's theta - and only then overwrites theta_ (particleIBMIO.C:196-204): until the first update() after the restart
  // the rotation matrix (and, for ellipses, the point mesh and the cell lists) are those of the initial orientation.  So the
  // lists are built once more with the dictionary angles in place, and the restored angles go back afterwards.
  if (c->nInst > 0) {
    const int m = c->nInst;
    std::vector<int> idx(m);
    std::vector<double> th(3 * (size_t)m), tmp(m);
    cudaMemcpy(idx.data(), c->in.index, m * sizeof(int), cudaMemcpyDeviceToHost);
    for (int q = 0; q < 3; ++q) {
      cudaMemcpy(th.data() + (size_t)q * m, c->in.th[q], m * sizeof(double), cudaMemcpyDeviceToHost);
      for (int i = 0; i < m; ++i) tmp[i] = c->orig[idx[i]].theta[q];
      cudaMemcpy(c->in.th[q], tmp.data(), m * sizeof(double), cudaMemcpyHostToDevice);
    }
    if ((rc = ctx_update(c))) return rc;
    if (c->nInst != m) return ctx_fail(c, IBM_E_UNSUPPORTED, "ibm_restore_particles: an instance is on the mesh with its restored angle and off it with the dictionary's");
    for (int q = 0; q < 3; ++q) cudaMemcpy(c->in.th[q], th.data() + (size_t)q * m, m * sizeof(double), cudaMemcpyHostToDevice);
  }
  // maxCopies_[index] = 0 for every surviving instance (particleCloud.C:1127-1131), -1 for indices without one
  std::fill(c->maxCopies.begin(), c->maxCopies.end(), -1);
  if (c->nInst > 0) {
    std::vector<int> idx(c->nInst);
    cudaMemcpy(idx.data(), c->in.index, c->nInst * sizeof(int), cudaMemcpyDeviceToHost);
    for (int i : idx) c->maxCopies[i] = 0;
  }
  motion_reset_copies(c);
  return IBM_OK;
}

int ibm_set_force_torque(ibm_ctx* c, const double* FT) {
  if (!c || !FT) return IBM_E_ARG;
  cudaSetDevice(c->device);
  const int n = c->nInst;
  std::vector<double> dv(n);
  for (int q = 0; q < 6; ++q) {
    for (int i = 0; i < n; ++i) dv[i] = FT[6 * (size_t)i + q];
    IBM_CUDA(c, cudaMemcpy(q < 3 ? c->in.F[q] : c->in.T[q - 3], dv.data(), n * sizeof(double), cudaMemcpyHostToDevice));
  }
  return IBM_OK;
}

int ibm_cd(ibm_ctx* c, double* out) {
  if (!c || !out) return IBM_E_ARG;
  cudaSetDevice(c->device);
  const int n = c->nInst;
  if (n == 0) return IBM_OK;
  IBM_CUDA(c, cudaStreamSynchronize(c->stream));
  std::vector<double> F(3 * (size_t)n);
  std::vector<int> cls(n);
  for (int q = 0; q < 3; ++q) IBM_CUDA(c, cudaMemcpy(F.data() + (size_t)q * n, c->in.F[q], n * sizeof(double), cudaMemcpyDeviceToHost));
  IBM_CUDA(c, cudaMemcpy(cls.data(), c->in.cls, n * sizeof(int), cudaMemcpyDeviceToHost));
  // particleIBM::Cd  particleIBM.C:583-600 (O(nInstances) scalar post-processing of the device result)
  const double* U = c->cfg.UInf;
  double magU = std::sqrt(U[0] * U[0] + U[1] * U[1] + U[2] * U[2]);
  double nU[3] = {U[0] / magU, U[1] / magU, U[2] / magU};
  double magSqrU = U[0] * U[0] + U[1] * U[1] + U[2] * U[2];
  for (int i = 0; i < n; ++i) {
    double dotF = nU[0] * F[i] + nU[1] * F[(size_t)n + i] + nU[2] * F[2 * (size_t)n + i];
    out[i] = dotF * 2.0 / (c->cfg.rhoRef * magSqrU * c->classes[cls[i]].dev.A);
  }
  return IBM_OK;
}

int ibm_max_co_num(ibm_ctx* c, double* out) {
  if (!c || !out) return IBM_E_ARG;
  *out = 0.0;  // particleCloud.C:1145-1156: the expression result is discarded, maxCo stays 0 (App. B #9)
  return IBM_OK;
}

int ibm_particle_mass(ibm_ctx* c, int inst, double* mass, double* inertia) {
  if (!c || inst < 0 || inst >= c->nInst) return IBM_E_ARG;
  cudaSetDevice(c->device);
  int cls;
  double rho;
  IBM_CUDA(c, cudaMemcpy(&cls, c->in.cls + inst, sizeof(int), cudaMemcpyDeviceToHost));
  IBM_CUDA(c, cudaMemcpy(&rho, c->in.rho + inst, sizeof(double), cudaMemcpyDeviceToHost));
  if (mass) *mass = c->classes[cls].dev.V * rho;
  if (inertia) *inertia = c->classes[cls].dev.I;
  return IBM_OK;
}

int ibm_substep_schedule(double dt, double* dts, int maxN) {
  // literal loop of particleCloud.C:1247-1275 (App. B #22)
  int n = 0;
  double localDt = dt / 100.0, localTime = localDt;
  while (localTime < dt) {
    if (dts && n < maxN) dts[n] = localDt;
    ++n;
    if (localTime + localDt > dt) localDt = dt - localTime;
    else localTime += localDt;
  }
  return n;
}

// ---- parity taps --------------------------------------------------------------------------------
int ibm_get_shell(ibm_ctx* c, int inst, int* cells, int* jk, int* Is, int* Os, double* w, double* W) {
  if (!c || inst < 0 || inst >= c->nInst) return ctx_fail(c, IBM_E_ARG, "ibm_get_shell: bad instance");
  cudaSetDevice(c->device);
  int n = 0;
  IBM_CUDA(c, cudaStreamSynchronize(c->stream));
  IBM_CUDA(c, cudaMemcpy(&n, c->in.nShell + inst, sizeof(int), cudaMemcpyDeviceToHost));
  if (n == 0) return IBM_OK;
  int *dI = nullptr;
  double* dD = nullptr;
  IBM_CUDA(c, cudaMalloc(&dI, (size_t)n * 15 * sizeof(int)));
  IBM_CUDA(c, cudaMalloc(&dD, (size_t)n * 9 * sizeof(double)));
  launch_tap_shell(c->stream, c->L, c->dClasses, c->in, inst, c->dShellCell, c->dShellJK, dI, dI + n, dI + 3 * (size_t)n, dI + 11 * (size_t)n, dD, dD + 8 * (size_t)n);
  std::vector<int> hI((size_t)n * 15);
  std::vector<double> hD((size_t)n * 9);
  cudaMemcpyAsync(hI.data(), dI, hI.size() * sizeof(int), cudaMemcpyDeviceToHost, c->stream);
  cudaMemcpyAsync(hD.data(), dD, hD.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
  cudaError_t e = cudaStreamSynchronize(c->stream);
  cudaFree(dI); cudaFree(dD);
  IBM_CUDA(c, e);
  // the reference lists shell cells in ascending cell id (its scan order, sphere.C:208)
  std::vector<int> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return hI[a] < hI[b]; });
  for (int o = 0; o < n; ++o) {
    int s = order[o];
    if (cells) cells[o] = hI[s];
    if (jk) { jk[2 * o] = hI[n + 2 * (size_t)s]; jk[2 * o + 1] = hI[n + 2 * (size_t)s + 1]; }
    for (int q = 0; q < 8; ++q) {
      if (Is) Is[8 * (size_t)o + q] = hI[3 * (size_t)n + 8 * (size_t)s + q];
      if (w) w[8 * (size_t)o + q] = hD[8 * (size_t)s + q];
    }
    for (int q = 0; q < 4; ++q) if (Os) Os[4 * (size_t)o + q] = hI[11 * (size_t)n + 4 * (size_t)s + q];
    if (W) W[o] = hD[8 * (size_t)n + s];
  }
  return IBM_OK;
}

int ibm_get_points(ibm_ctx* c, int inst, double* pts, int* cell, int* nFaces, int* otherCell, double* w, double* W) {
  if (!c || inst < 0 || inst >= c->nInst) return ctx_fail(c, IBM_E_ARG, "ibm_get_points: bad instance");
  cudaSetDevice(c->device);
  int cls = 0;
  IBM_CUDA(c, cudaStreamSynchronize(c->stream));
  IBM_CUDA(c, cudaMemcpy(&cls, c->in.cls + inst, sizeof(int), cudaMemcpyDeviceToHost));
  const size_t N = (size_t)c->classes[cls].dev.N;
  int* dI = nullptr;
  double* dD = nullptr;
  IBM_CUDA(c, cudaMalloc(&dI, N * 8 * sizeof(int)));
  IBM_CUDA(c, cudaMalloc(&dD, N * 10 * sizeof(double)));
  launch_tap_points(c->stream, c->L, c->dClasses, c->in, inst, dD, dI, dI + N, dI + 2 * N, dD + 3 * N, dD + 9 * N);
  std::vector<int> hI(N * 8);
  std::vector<double> hD(N * 10);
  cudaMemcpyAsync(hI.data(), dI, hI.size() * sizeof(int), cudaMemcpyDeviceToHost, c->stream);
  cudaMemcpyAsync(hD.data(), dD, hD.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
  cudaError_t e = cudaStreamSynchronize(c->stream);
  cudaFree(dI); cudaFree(dD);
  IBM_CUDA(c, e);
  if (pts) std::memcpy(pts, hD.data(), 3 * N * sizeof(double));
  if (cell) std::memcpy(cell, hI.data(), N * sizeof(int));
  if (nFaces) std::memcpy(nFaces, hI.data() + N, N * sizeof(int));
  if (otherCell) std::memcpy(otherCell, hI.data() + 2 * N, 6 * N * sizeof(int));
  if (w) std::memcpy(w, hD.data() + 3 * N, 6 * N * sizeof(double));
  if (W) std::memcpy(W, hD.data() + 9 * N, N * sizeof(double));
  return IBM_OK;
}

int ibm_get_sf(ibm_ctx* c, int inst, double* Sf) {
  if (!c || !Sf || inst < 0 || inst >= c->nInst) return IBM_E_ARG;
  cudaSetDevice(c->device);
  int cls = 0;
  IBM_CUDA(c, cudaMemcpy(&cls, c->in.cls + inst, sizeof(int), cudaMemcpyDeviceToHost));
  const ShapeClassHost& h = c->classes[cls];
  IBM_CUDA(c, cudaMemcpy(Sf, h.dev.Sf, h.Sf.size() * sizeof(double), cudaMemcpyDeviceToHost));
  return IBM_OK;
}

int ibm_get_indicator(ibm_ctx* c, double* I) {
  if (!c || !I) return IBM_E_ARG;
  cudaSetDevice(c->device);
  if (c->cap == 0) return ctx_fail(c, IBM_E_STATE, "ibm_get_indicator: no particles set");
  int rc = ensure_field(c, &c->dI, 1);
  if (rc) return rc;
  if (c->indicatorStale) {
    ctx_tic(c, T_UPDATE);
    IBM_CUDA(c, cudaMemsetAsync(c->dI, 0, (size_t)c->L.nCells * sizeof(double), c->stream));
    launch_k1_indicator(c->stream, c->L, c->dClasses, c->in, c->nInst, c->dI);
    c->nLaunch += 2;
    ctx_toc(c);
    c->indicatorStale = false;
  }
  rc = d2h_field(c, c->dI, 1, I);
  if (rc) return rc;
  IBM_CUDA(c, cudaStreamSynchronize(c->stream));
  return IBM_OK;
}

// EXTENSION (no reference counterpart): fractional solid volume fraction, nSub^D sample points per cell
int ibm_get_solid_fraction(ibm_ctx* c, int nSub, double* alpha) {
  if (!c || !alpha) return IBM_E_ARG;
  if (nSub < 1 || nSub > 32) return ctx_fail(c, IBM_E_ARG, "ibm_get_solid_fraction: nSub must be 1..32");
  cudaSetDevice(c->device);
  if (c->cap == 0) return ctx_fail(c, IBM_E_STATE, "ibm_get_solid_fraction: no particles set");
  double* dA = nullptr;
  IBM_CUDA(c, cudaMalloc(&dA, (size_t)c->L.nCells * sizeof(double)));
  launch_solid_fraction(c->stream, c->L, c->dClasses, c->in, c->nInst, nSub, dA);
  c->nLaunch += 2;
  int rc = d2h_field(c, dA, 1, alpha);
  cudaError_t e = cudaStreamSynchronize(c->stream);
  cudaFree(dA);
  if (rc) return rc;
  IBM_CUDA(c, e);
  return IBM_OK;
}

// ---- device-resident helpers ---------------------------------------------------------------------
int ibm_device_field_bytes(ibm_ctx* c, int nComp, size_t* bytes) {
  if (!c || !bytes || nComp <= 0) return IBM_E_ARG;
  *bytes = (size_t)nComp * c->L.nCells * sizeof(double);
  return IBM_OK;
}
int ibm_upload_field(ibm_ctx* c, const double* hostAoS, int nComp, double* dSoA) {
  if (!c || !hostAoS || !dSoA) return IBM_E_ARG;
  cudaSetDevice(c->device);
  return h2d_field(c, hostAoS, nComp, dSoA);
}
int ibm_download_field(ibm_ctx* c, const double* dSoA, int nComp, double* hostAoS) {
  if (!c || !hostAoS) return IBM_E_ARG;
  cudaSetDevice(c->device);
  const double* src = dSoA ? dSoA : c->dF;  // NULL = the persistent forcing field
  if (!src) return ctx_fail(c, IBM_E_STATE, "ibm_download_field: no forcing field yet");
  int rc = d2h_field(c, src, nComp, hostAoS);
  if (rc) return rc;
  IBM_CUDA(c, cudaStreamSynchronize(c->stream));
  return IBM_OK;
}
// cudaHostRegister of a caller-owned array that persists across calls (shim: U, U.oldTime(), p): its transfers then run at the
// full DMA rate instead of through the staging chunks.  Unregister before the array is freed.
int ibm_host_register(ibm_ctx* c, const void* ptr, size_t bytes) {
  if (!c || !ptr || bytes == 0) return IBM_E_ARG;
  cudaSetDevice(c->device);
  if (host_is_pinned(ptr)) return IBM_OK;
  cudaError_t e = cudaHostRegister(const_cast<void*>(ptr), bytes, cudaHostRegisterDefault);
  if (e != cudaSuccess) { cudaGetLastError(); return ctx_fail(c, IBM_E_CUDA, std::string("cudaHostRegister: ") + cudaGetErrorString(e)); }
  return IBM_OK;
}
int ibm_host_unregister(ibm_ctx* c, const void* ptr) {
  if (!c || !ptr) return IBM_E_ARG;
  cudaSetDevice(c->device);
  cudaError_t e = cudaHostUnregister(const_cast<void*>(ptr));
  if (e != cudaSuccess) { cudaGetLastError(); return ctx_fail(c, IBM_E_CUDA, std::string("cudaHostUnregister: ") + cudaGetErrorString(e)); }
  return IBM_OK;
}

int ibm_sync(ibm_ctx* c) {
  if (!c) return IBM_E_ARG;
  cudaSetDevice(c->device);
  IBM_CUDA(c, cudaStreamSynchronize(c->stream));
  return IBM_OK;
}
int ibm_set_kernel_variant(ibm_ctx* c, int fv, int sv) {
  if (!c) return IBM_E_ARG;
  c->forcingVariant = fv;
  c->stressVariant = sv;
  if (c->cap != 0) return ctx_update(c);  // the work-item table depends on the forcing kernel generation
  return IBM_OK;
}

int ibm_get_counter(ibm_ctx* c, const char* name, long long* value) {
  if (!c || !name || !value) return IBM_E_ARG;
  if (std::strcmp(name, "forcing_variant") == 0) { *value = c->forcingVariant; return IBM_OK; }
  if (std::strcmp(name, "stress_variant") == 0) { *value = c->stressVariant; return IBM_OK; }
  if (std::strcmp(name, "mark_variant") == 0) { *value = c->markVariant; return IBM_OK; }
  if (std::strcmp(name, "h2d_pinned_bytes") == 0) { *value = c->h2dPinnedBytes; return IBM_OK; }   // since the last ibm_timings_reset
  if (std::strcmp(name, "h2d_staged_bytes") == 0) { *value = c->h2dStagedBytes; return IBM_OK; }
  if (std::strcmp(name, "v7_ok") == 0) { *value = c->v7ok ? 1 : 0; return IBM_OK; }
  if (std::strcmp(name, "v7_slow_count") == 0) {  // instances the last staged launch handed to the direct-gather kernel
    int n = 0;
    cudaStreamSynchronize(c->stream);
    cudaMemcpy(&n, c->dScal + 12, sizeof(int), cudaMemcpyDeviceToHost);
    *value = n;
    return IBM_OK;
  }
  if (std::strcmp(name, "v7_smem_forcing") == 0) { *value = c->v7ok ? c->v7capF.planeCap * 8LL + c->v7capF.colCap * 48LL : 0; return IBM_OK; }
  return motion_counter(c, name, value) ? IBM_OK : ctx_fail(c, IBM_E_ARG, std::string("ibm_get_counter: unknown counter ") + name);
}

int ibm_set_option(ibm_ctx* c, const char* name, int value) {
  if (!c || !name) return IBM_E_ARG;
  if (std::strcmp(name, "fused_loop5") == 0) { c->fusedLoop5 = value != 0; return IBM_OK; }
  if (std::strcmp(name, "graph_loop5") == 0) { c->graphLoop5 = value != 0; return IBM_OK; }
  if (std::strcmp(name, "chunk256") == 0) { c->chunk256 = value != 0; return c->cap != 0 ? ctx_update(c) : IBM_OK; }
  if (std::strcmp(name, "active_loop5") == 0) { c->activeLoop5 = value != 0; return IBM_OK; }
  if (std::strcmp(name, "cluster_loop5") == 0) { c->clusterLoop5 = value != 0; return IBM_OK; }
  if (std::strcmp(name, "owned_loop5") == 0) { c->ownedLoop5 = value != 0; return IBM_OK; }
  if (std::strcmp(name, "cluster_one_block_cc") == 0) { c->clusterOneBlockCC = value != 0; return IBM_OK; }
  if (std::strcmp(name, "async_prepare") == 0) { c->asyncPrepare = value != 0; return IBM_OK; }
  if (std::strcmp(name, "cluster_general_only") == 0) { c->clusterGeneralOnly = value != 0; return IBM_OK; }
  if (std::strcmp(name, "active_loop5_safety_ppm") == 0) { c->activeSafetyPpm = value; return IBM_OK; }
  if (std::strcmp(name, "fused_active_loop5") == 0) { c->fusedActive = value != 0; return IBM_OK; }
  if (std::strcmp(name, "active_loop5_check_ppm") == 0) { c->activeCheckPpm = value; return IBM_OK; }
  if (std::strcmp(name, "v7_plane_cap") == 0) {  // test hook: same for the staged planes (forcing and stress)
    if (value < 0) return ctx_fail(c, IBM_E_ARG, "ibm_set_option: v7_plane_cap >= 0");
    c->v7capF.planeCap = std::min(c->v7capF.planeCap, value);
    c->v7capS.planeCap = std::min(c->v7capS.planeCap, value);
    return IBM_OK;
  }
  if (std::strcmp(name, "mark_variant") == 0) {
    if (value != 0 && value != 1) return ctx_fail(c, IBM_E_ARG, "ibm_set_option: mark_variant is 0 or 1");
    c->markVariant = value;
    return IBM_OK;
  }
  if (std::strcmp(name, "collision_model") == 0) {
    if (value != 0 && value != 1) return ctx_fail(c, IBM_E_ARG, "ibm_set_option: collision_model is 0 (reference impulse) or 1 (soft sphere)");
    c->collisionModel = value;
    motion_invalidate_graph(c);
    return IBM_OK;
  }
  return ctx_fail(c, IBM_E_ARG, std::string("ibm_set_option: unknown option ") + name);
}

int ibm_set_real_option(ibm_ctx* c, const char* name, double value) {
  if (!c || !name) return IBM_E_ARG;
  double* slot = nullptr;
  if (std::strcmp(name, "soft_sphere_kn") == 0) slot = &c->softKn;
  else if (std::strcmp(name, "soft_sphere_gamma_n") == 0) slot = &c->softGammaN;
  else if (std::strcmp(name, "soft_sphere_gamma_t") == 0) slot = &c->softGammaT;
  else if (std::strcmp(name, "soft_sphere_mu") == 0) slot = &c->softMu;
  if (!slot) return ctx_fail(c, IBM_E_ARG, std::string("ibm_set_real_option: unknown option ") + name);
  if (!(value >= 0.0)) return ctx_fail(c, IBM_E_ARG, std::string("ibm_set_real_option: ") + name + " must be >= 0");
  *slot = value;
  motion_invalidate_graph(c);
  return IBM_OK;
}

int ibm_timings(ibm_ctx* c, ibm_timing* out) {
  if (!c || !out) return IBM_E_ARG;
  cudaSetDevice(c->device);
  resolve_events(c);
  out->update_ms = c->accMs[T_UPDATE]; out->forcing_ms = c->accMs[T_FORCING]; out->stress_ms = c->accMs[T_STRESS];
  out->advance_ms = c->accMs[T_ADVANCE]; out->h2d_ms = c->accMs[T_H2D]; out->d2h_ms = c->accMs[T_D2H];
  out->layout_ms = c->accMs[T_LAYOUT];
  out->h2d_bytes = c->h2dBytes; out->d2h_bytes = c->d2hBytes; out->n_kernel_launches = c->nLaunch;
  out->n_forcing_calls = c->nForcingCalls; out->n_updates = c->nUpdates;
  return IBM_OK;
}
int ibm_timings_reset(ibm_ctx* c) {
  if (!c) return IBM_E_ARG;
  resolve_events(c);
  for (double& x : c->accMs) x = 0;
  c->h2dBytes = c->d2hBytes = c->nLaunch = c->nForcingCalls = c->nUpdates = 0;
  c->h2dPinnedBytes = c->h2dStagedBytes = 0;
  return IBM_OK;
}

}  // extern "C"
