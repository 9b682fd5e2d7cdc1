// gi_kernels.inl -- ReSTIR GI: ReSTIRIndirectKernel (restir.cu:242-416) for sm_100a.  Included by kernels.cu (namespace rs), after the
// ray / material / light helpers it shares with the direct-illumination kernels.
//
//   k_restir_indirect        one thread per pixel.  The jittered primary rays of an 8x4 tile walk the traced tree as one packet
//                            (packetWalk, as in k_ptdirect); every bounce ray then walks it per lane (traceClosestFast: the same
//                            leaf-box / near-tie criterion, so the reported hit is the reference's), next-event shadow rays use the
//                            any-hit walk.  A pixel with an undecided ray (more mutually near hits than the tie store holds) writes
//                            nothing and is queued.
//   k_gi_primary / k_gi_bounce / k_gi_resolve   the same frame as a wavefront (rstr_gi_set_pipeline): one launch per iteration of the path
//                            loop over the compacted live paths; bit-identical to k_restir_indirect (see "the staged form" below)
//   k_gi_head / k_gi_walk_shadow / k_gi_walk_closest / k_gi_tail   the wavefront with the walks of an iteration as persistent kernels over
//                            ray lists, lanes refilled from the list as they finish (see "the ray-queue form"); the default on real scenes
//   k_restir_indirect_fix    the queued pixels again with the reference-order walk of the reference tree
//   k_restir_indirect_exact  every ray with the reference-order walk (RS_TRAVERSAL_EXACT: validation mode)
//   k_export_gi              device reservoirs -> Reservoir<IndirectLiSample> (68 B: Lo xv nv xs ns, numSamples, weight; restir.h:13-27)
//
// The reference ships this kernel disabled (commented out at main.cpp:168, Settings::traceDepth = 0 in common.cpp:3); the trace depth
// and the temporal switch are arguments here.  One deviation, shared with the oracle: the reference jumps to WriteSample past the
// initialisation of primMaterial / primWo / primSampleDelta when the jittered ray leaves the scene or hits an emitter and then shades a
// history sample with them (undefined behaviour); such a pixel writes indirect = 0.
//
// Arithmetic: every expression below keeps the operation order of the reference's (cited per function); the BSDF sampling goes through
// sinf / cosf, so against a CPU build of the same expressions (glibc) the sampled directions agree to an ulp, not to the bit.
//
// Per pixel and frame: 4 + (depth - 1) * 4 + depth * 3 + 2 LCG draws at most, depth closest-hit rays + (depth - 1) shadow rays,
// 68 B of reservoir written (64-B record + 4-B plane), 68 B of history gathered at the reprojected pixel, 24 B radiance read-modify-write.

struct GISample { f3 Lo, xv, nv, xs, ns; };                                          // IndirectLiSample, restir.h:13-27
struct GIResv { GISample s; int M; float w; };                                       // Reservoir<IndirectLiSample>, restir.h:29-117
RS_D GISample giEmptySample() { GISample s; s.Lo = s.xv = s.nv = s.xs = s.ns = mk3(0.f); return s; }
RS_D bool giResvInvalid(const GIResv& r) { return isNanOrInf(r.w) || r.w < 0.f; }    // restir.h:51

// HBM form: one 64-byte record {Lo, weight} {xv, numSamples} {nv, ns.x} {xs, ns.y} (4 x LDG / STG.128 on one 64-B aligned line)
// plus ns.z in a separate 4-byte plane: 68 B per pixel, the reference's size, without a record straddling sectors
RS_D GIResv giLoad(const float4* rec, const float* nsz, size_t i) {
    const float4 a = __ldg(rec + 4 * i), b = __ldg(rec + 4 * i + 1), c = __ldg(rec + 4 * i + 2), d = __ldg(rec + 4 * i + 3);
    GIResv r;
    r.s.Lo = mk3(a.x, a.y, a.z); r.w = a.w;
    r.s.xv = mk3(b.x, b.y, b.z); r.M = __float_as_int(b.w);
    r.s.nv = mk3(c.x, c.y, c.z);
    r.s.xs = mk3(d.x, d.y, d.z);
    r.s.ns = mk3(c.w, d.w, __ldg(nsz + i));
    return r;
}
RS_D void giStore(float4* rec, float* nsz, size_t i, const GIResv& r) {
    rec[4 * i] = make_float4(r.s.Lo.x, r.s.Lo.y, r.s.Lo.z, r.w);
    rec[4 * i + 1] = make_float4(r.s.xv.x, r.s.xv.y, r.s.xv.z, __int_as_float(r.M));
    rec[4 * i + 2] = make_float4(r.s.nv.x, r.s.nv.y, r.s.nv.z, r.s.ns.x);
    rec[4 * i + 3] = make_float4(r.s.xs.x, r.s.xs.y, r.s.xs.z, r.s.ns.y);
    nsz[i] = r.s.ns.z;
}

// ------------------------------------------------------------------------------------------------ closest hit of one ray, traced tree
// The walk of traceOccludedFast with the running result of the packet walk (PRay: best hit, near ties, limit) kept per lane.
// false = undecided.
RS_D bool traceClosestFast(const DevScene& s, f3 o, f3 d, Stack& stack, const TieStore& ts, Hit& h) {
    // a NaN direction (sampleHemisphereCosine with r.x == 1: sqrt of a negative rounding residue) fails every box test of the reference
    // (bvh.h:85-157: all comparisons false) -> miss; the slab test's fminf / fmaxf would drop the NaNs and walk the whole tree instead
    if (isnan(d.x) || isnan(d.y) || isnan(d.z) || isnan(o.x) || isnan(o.y) || isnan(o.z)) {
        h.t = FLT_MAX; h.prim = -1; h.bx = 0.f; h.by = 0.f;
        return true;
    }
    PRay p = prayBegin(o, d, true, ts);
    float t0;
    if (slabHitP(p, s.fastRootMin[0], s.fastRootMin[1], s.fastRootMin[2], s.fastRootMax[0], s.fastRootMax[1], s.fastRootMax[2], t0)) {
        int sp = 0;
        int cur = s.fastRoot;
        for (;;) {
            bool popped = true;
            while (cur >= 0) {
                const float4* np = s.fastNodes + 4 * (size_t)cur;
                const F8 nA = ldg256(np), nB = ldg256(np + 2);
                const float4 a = nA.lo, b = nA.hi, c = nB.lo;
                const int2 l = make_int2(__float_as_int(nB.hi.x), __float_as_int(nB.hi.y));
                float tL, tR;
                const bool hL = slabHitP(p, a.x, a.y, a.z, a.w, b.x, b.y, tL);
                const bool hR = slabHitP(p, b.z, b.w, c.x, c.y, c.z, c.w, tR);
                if (hL && hR) {
                    const bool leftNear = tL <= tR;
                    stack.push(sp, leftNear ? l.y : l.x, leftNear ? tR : tL); sp++;
                    cur = leftNear ? l.x : l.y;
                } else if (hL) cur = l.x;
                else if (hR) cur = l.y;
                else {
                    popped = false;
                    while (sp > 0) {
                        --sp;
                        if (stack.t(sp) <= p.limit) { cur = stack.ref(sp); popped = true; break; }
                    }
                    if (!popped) break;
                }
            }
            if (!popped) break;
            const int first = cur & 0x07ffffff, count = ((cur >> 27) & 7) + 1;
            for (int i = 0; i < count; i++) {
                const Tri t = loadTriFast(s, first + i);
                prayOffer(p, o, d, ts, t, first + i);
            }
            popped = false;
            while (sp > 0) {
                --sp;
                if (stack.t(sp) <= p.limit) { cur = stack.ref(sp); popped = true; break; }
            }
            if (!popped) break;
        }
    }
    return prayResolve(s, o, d, p, ts, h);
}

template <bool EXACT>
RS_D bool giClosest(const DevScene& s, const GIDev& g, f3 o, f3 d, Stack& stack, const TieStore& ts, Hit& h) {
    if (EXACT || g.bounceWalk == RS_TRAVERSAL_EXACT) {
        const RayT r = makeRayT(o, d);
        traceClosestExact(s, r, h, stack);
        return true;
    }
    return traceClosestFast(s, o, d, stack, ts, h);
}

// ------------------------------------------------------------------------------------------------ surface at a hit
// DevScene::getIntersecGeomInfo + getTexturedMaterialAndSurface (scene.h:135-151, 78-99), as in ptdirectAfterHit
struct GISurf { f3 pos, nrm, v0, v1, v2; Surf m; float ior; };
RS_D void giSurface(const DevScene& s, const Hit& h, GISurf& g) {
    const Tri t = loadTri(s, h.prim);
    const float4* np = s.triNorm + 3 * (size_t)h.prim;
    const float4 na4 = __ldg(np), nb4 = __ldg(np + 1), nc4 = __ldg(np + 2);
    const f3 na = mk3(na4.x, na4.y, na4.z), nb = mk3(na4.w, nb4.x, nb4.y), nc = mk3(nb4.z, nb4.w, nc4.x);
    const float bz = 1.f - h.bx - h.by;
    g.pos = t.v1 * h.bx + t.v2 * h.by + t.v0 * bz;
    g.nrm = normalize(nb * h.bx + nc * h.by + na * bz);
    g.v0 = t.v0; g.v1 = t.v1; g.v2 = t.v2;
    g.m = texturedMaterial(s, t.matId, h.prim, h.bx, h.by, g.nrm);
    g.ior = __ldg(&s.materials[t.matId].ior);
}

// ------------------------------------------------------------------------------------------------ Material::sample / pdf (material.h)
RS_D f3 giMatVec(f3 c0, f3 c1, f3 c2, f3 v) {                                        // type_mat3x3.inl operator*(mat3, vec3)
    return mk3(c0.x * v.x + c1.x * v.y + c2.x * v.z, c0.y * v.x + c1.y * v.y + c2.y * v.z, c0.z * v.x + c1.z * v.y + c2.z * v.z);
}
RS_D f3 giReflect(f3 I, f3 N) { return I - N * dot(N, I) * 2.f; }                    // func_geometric.inl:176
RS_D void giConcentricDisk(float rx, float ry, float& px, float& py) {               // mathUtil.h:128-132
    const float r = sqrtf(rx), theta = ry * RS_PI * 2.0f;
    px = cosf(theta) * r; py = sinf(theta) * r;
}
RS_D f3 giSampleHemisphereCosine(f3 n, float rx, float ry) {                         // mathUtil.h:157-161
    float dx, dy;
    giConcentricDisk(rx, ry, dx, dy);
    const float z = sqrtf(1.f - (dx * dx + dy * dy));
    return localToWorld(n, mk3(dx, dy, z));
}
RS_D float giFresnel(float cosIn, float ior) {                                       // material.h:43-60 (the #if tests a misspelt macro: the exact branch)
    if (cosIn < 0) { ior = 1.f / ior; cosIn = -cosIn; }
    const float sinIn = sqrtf(1.f - cosIn * cosIn);
    const float sinTr = sinIn / ior;
    if (sinTr >= 1.f) return 1.f;
    const float cosTr = sqrtf(1.f - sinTr * sinTr);
    const float a = (cosIn - ior * cosTr) / (cosIn + ior * cosTr), b = (ior * cosIn - cosTr) / (ior * cosIn + cosTr);
    return (a * a + b * b) * .5f;
}
RS_D bool giRefract(f3 n, f3 wi, float ior, f3& wt) {                                // mathUtil.h:163-180
    const float cosIn = dot(n, wi);
    if (cosIn < 0) ior = 1.f / ior;
    const float sin2In = gmax(0.f, 1.f - cosIn * cosIn);
    const float sin2Tr = sin2In / (ior * ior);
    if (sin2Tr >= 1.f) return false;
    float cosTr = sqrtf(1.f - sin2Tr);
    if (cosIn < 0) cosTr = -cosTr;
    wt = normalize(-wi / ior + n * (cosIn / ior - cosTr));
    return true;
}
RS_D float giGTR2Pdf(f3 n, f3 m, f3 wo, float alpha) {                               // material.h:82-85
    return GTR2Distrib(dot(n, m), alpha) * schlickG(dot(n, wo), alpha) * absDot(m, wo) / absDot(n, wo);
}
// GGX visible-normal sampling, material.h:93-112 (localRefMatrix mathUtil.h:146-151, glm::inverse(mat3) type_mat3x3.inl:37-57)
__device__ __noinline__ f3 giGTR2Sample(f3 n, f3 wo, float alpha, float rx, float ry) {
    f3 t0 = (fabsf(n.y) > 0.9999f) ? mk3(0.f, 0.f, 1.f) : mk3(0.f, 1.f, 0.f);
    const f3 b0 = normalize(cross(n, t0));
    t0 = cross(b0, n);
    const f3 m0 = t0, m1 = b0, m2 = n;                                                // columns
    const float m00 = m0.x, m01 = m0.y, m02 = m0.z, m10 = m1.x, m11 = m1.y, m12 = m1.z, m20 = m2.x, m21 = m2.y, m22 = m2.z;
    const float ood = 1.f / (+m00 * (m11 * m22 - m21 * m12) - m10 * (m01 * m22 - m21 * m02) + m20 * (m01 * m12 - m11 * m02));
    f3 i0, i1, i2;                                                                    // columns of the inverse
    i0.x = +(m11 * m22 - m21 * m12) * ood; i1.x = -(m10 * m22 - m20 * m12) * ood; i2.x = +(m10 * m21 - m20 * m11) * ood;
    i0.y = -(m01 * m22 - m21 * m02) * ood; i1.y = +(m00 * m22 - m20 * m02) * ood; i2.y = -(m00 * m21 - m20 * m01) * ood;
    i0.z = +(m01 * m12 - m11 * m02) * ood; i1.z = -(m00 * m12 - m10 * m02) * ood; i2.z = +(m00 * m11 - m10 * m01) * ood;
    const f3 vh = normalize(giMatVec(i0, i1, i2, wo) * mk3(alpha, alpha, 1.f));
    const float lenSq = vh.x * vh.x + vh.y * vh.y;
    const f3 t = lenSq > 0.f ? mk3(-vh.y, vh.x, 0.f) / sqrtf(lenSq) : mk3(1.f, 0.f, 0.f);
    const f3 b = cross(vh, t);
    float px, py;
    giConcentricDisk(rx, ry, px, py);
    const float sgn = 0.5f * (vh.z + 1.f);
    py = (1.f - sgn) * sqrtf(1.f - px * px) + sgn * py;
    f3 h = t * px + b * py + vh * sqrtf(gmax(0.f, 1.f - (px * px + py * py)));
    h = mk3(h.x * alpha, h.y * alpha, gmax(0.f, h.z));
    return normalize(giMatVec(m0, m1, m2, h));
}
RS_D float giMaterialPdf(const Surf& m, f3 n, f3 wo, f3 wi) {                        // material.h:230-240
    if (m.type == 0) return satDot(n, wi) * 1.f / RS_PI;
    if (m.type == 1) {
        const f3 h = normalize(wo + wi);
        return mixf(satDot(n, wi) * 1.f / RS_PI, giGTR2Pdf(n, h, wo, m.roughness * m.roughness) / (4.f * absDot(h, wo)), 1.f / (2.f - m.metallic));
    }
    return 0.f;
}
RS_D f3 giBSDF(const Surf& m, f3 n, f3 wo, f3 wi) { return materialBSDF(m.type, m.metallic, m.roughness, m.baseColor, diffuseTerm(m.baseColor), n, wo, wi); }

enum { GI_BS_INVALID = 0, GI_BS_SMOOTH = 1, GI_BS_SPECULAR = 2 };                    // what the kernel reads of BSDFSampleType (material.h:15-24)
struct GIBSample { f3 dir, bsdf; float pdf; int kind; };
RS_D void giMaterialSample(const Surf& m, float ior, f3 n, f3 wo, float rx, float ry, float rz, GIBSample& s) {   // material.h:242-256
    s.dir = mk3(0.f); s.bsdf = mk3(0.f); s.pdf = 0.f; s.kind = GI_BS_INVALID;
    if (m.type == 0) {                                                                // lambertianSample :130-135
        s.dir = giSampleHemisphereCosine(n, rx, ry);
        s.bsdf = m.baseColor * 1.f / RS_PI;
        s.pdf = satDot(n, s.dir) * 1.f / RS_PI;
        s.kind = GI_BS_SMOOTH;
    } else if (m.type == 1) {                                                         // metallicWorkflowSample :197-216
        const float alpha = m.roughness * m.roughness;
        if (rz > (1.f / (2.f - m.metallic))) s.dir = giSampleHemisphereCosine(n, rx, ry);
        else {
            const f3 h = giGTR2Sample(n, wo, alpha, rx, ry);
            s.dir = -giReflect(wo, h);
        }
        if (!(dot(n, s.dir) < 0.f)) {
            s.bsdf = giBSDF(m, n, wo, s.dir);
            s.pdf = giMaterialPdf(m, n, wo, s.dir);
            s.kind = GI_BS_SMOOTH;
        }
    } else if (m.type == 2) {                                                         // dielectricSample :145-169
        const float pdfRefl = giFresnel(dot(n, wo), ior);
        s.bsdf = m.baseColor;
        if (rz < pdfRefl) {
            s.dir = giReflect(-wo, n);
            s.kind = GI_BS_SPECULAR;
            s.pdf = 1.f;
        } else {
            if (!giRefract(n, wo, ior, s.dir)) return;
            float eta = ior;
            if (dot(n, wo) < 0) eta = 1.f / eta;
            s.bsdf = s.bsdf / (eta * eta);
            s.kind = GI_BS_SPECULAR;
            s.pdf = 1.f;
        }
    }
}
RS_D float giPowerHeuristic(float f, float g) { const float f2 = f * f; return f2 / (f2 + g * g); }   // mathUtil.h:81-84

// ------------------------------------------------------------------------------------------------ DevScene::sampleDirectLight
// scene.h:427-459, :377-392 for the environment map, without the occlusion test: the pdf (<= 0 or NaN: no contribution whatever the test
// would say) and the point `target` the segment from pos has to reach.  The reference tests occlusion first and the facing of the light
// second; both only ever turn a contribution into none, so testing the facing first saves the ray without changing a bit.
RS_D float giLightSample(const DevScene& s, f3 pos, float c0, float c1, float c2, float c3, f3& Li, f3& wi, f3& target) {
    const int len = s.numLights;
    target = pos;
    if (len <= 0) return -1.f;
    const int pass = min(__float2int_rz((float)len * c0), len - 1);                   // sampler.h:203-207
    const float2 e = __ldg(s.alias + pass);
    const int lightId = (c1 < e.x) ? pass : __float_as_int(e.y);
    if (s.envTex >= 0 && lightId == len - 1) {
        envSample(s, c2, c3, Li, wi);
        target = pos + wi * 1e6f;
        return envPdf(s, Li);
    }
    const float4* lp = s.lights + 4 * (size_t)lightId;
    const float4 a = __ldg(lp), b = __ldg(lp + 1), c = __ldg(lp + 2), d4 = __ldg(lp + 3);
    const f3 v0 = mk3(a.x, a.y, a.z), v1 = mk3(a.w, b.x, b.y), v2 = mk3(b.z, b.w, c.x);
    const f3 n = mk3(c.y, c.z, c.w);
    const float sr = sqrtf(c3);                                                       // mathUtil.h:94-100 (ru = r.z, rv = r.w)
    const float u = 1.f - sr, v = c2 * sr;
    const f3 sampled = v1 * u + v2 * v + v0 * (1.f - u - v);
    target = sampled;
    const f3 pts = sampled - pos;
    if (dot(n, pts) > -1e-6f) return -1.f;                                            // SCENE_LIGHT_SINGLE_SIDED
    Li = mk3(d4.x, d4.y, d4.z);
    const float len2 = dot(pts, pts);
    wi = pts * (1.f / sqrtf(len2));
    return d4.w * len2 / fabsf(dot(n, wi));                                           // pdfAreaToSolidAngle (see sampleLight)
}

// findTemporalNeighbor<IndirectReservoir> (restir.cu:20-45): index of the history pixel in the planes, or -1
RS_D long long giFindTemporal(const FrameDev& f, size_t li) {
    const int primId = f.matId[0][li];
    const int lastIdx = __float_as_int(f.albedoMotion[li].w);
    if (lastIdx < 0 || primId <= -1) return -1;
    if (!rowResident(f, lastIdx / f.W)) return -1;                                    // GI runs on full frames: never taken
    const size_t lli = (size_t)lastIdx - (size_t)f.bufRow0 * f.W;
    if (f.matId[1][lli] != primId) return -1;
    const float4 g = f.geom[0][li], lg = f.geom[1][lli];
    if (absDot(mk3(g.x, g.y, g.z), mk3(lg.x, lg.y, lg.z)) < .9f || fabsf(lg.w - g.w) > g.w * .1f) return -1;
    return (long long)lli;
}

// ------------------------------------------------------------------------------------------------ the pixel
// restir.cu:263-415 in three pieces, so that the one-kernel form (giAfterHit) and the staged form (k_gi_primary / k_gi_bounce /
// k_gi_resolve) run the very same expressions: giHead = one iteration of the path loop up to its bounce ray (:283-318),
// giTail = the ray and what it hits (:320-375), giWriteSample = WriteSample (:380-415).
struct GIVertex { f3 pos, nrm, wo; Surf mat; float ior; };                           // a path vertex whose BSDF is sampled next
struct GIRay { f3 dir; float pdf; bool delta; };                                     // the sampled bounce: makeOffsetedRay(pos, dir)

// :284-286
RS_D void giFaceForward(GIVertex& v) { if (v.mat.type != 2 && dot(v.nrm, v.wo) < 0.f) v.nrm = -v.nrm; }
// :288-299: next-event estimation with MIS at a vertex after the first, in two steps: the light sample and what it would contribute
// (c) if the segment to `target` is free, then the occlusion test
struct GINee { f3 c, target; bool want; };
RS_D GINee giNextEventSample(const DevScene& s, const GIVertex& v, f3 throughput, Rng& rng) {
    const float c0 = rng.next(), c1 = rng.next(), c2 = rng.next(), c3 = rng.next();
    f3 radiance = mk3(0.f), wi = mk3(0.f);
    GINee n;
    n.c = mk3(0.f);
    const float lightPdf = giLightSample(s, v.pos, c0, c1, c2, c3, radiance, wi, n.target);
    n.want = lightPdf > 0.f;
    if (n.want) {
        const float bsdfPdf = giMaterialPdf(v.mat, v.nrm, v.wo, wi);
        n.c = throughput * giBSDF(v.mat, v.nrm, v.wo, wi) * radiance * satDot(v.nrm, wi) / lightPdf * giPowerHeuristic(lightPdf, bsdfPdf);
    }
    return n;
}
// false: undecided shadow ray (never with EXACT)
template <bool EXACT>
RS_D bool giNextEvent(const DevScene& s, const GIVertex& v, f3 throughput, f3& Lo, Rng& rng, Stack& stack) {
    const GINee n = giNextEventSample(s, v, throughput, rng);
    if (n.want) {
        const int occ = traceOccluded<EXACT>(s, v.pos, n.target, stack);
        if (occ < 0) return false;
        if (!occ) Lo = Lo + n.c;
    }
    return true;
}
// :301-318: the bounce direction.  true: ray sampled, false: the path ends here (the loop's break)
RS_D bool giSampleBounce(int depth, const GIVertex& v, f3& throughput, Rng& rng, GIRay& ray) {
    const float r0 = rng.next(), r1 = rng.next(), r2 = rng.next();
    GIBSample bs;
    giMaterialSample(v.mat, v.ior, v.nrm, v.wo, r0, r1, r2, bs);
    if (bs.kind == GI_BS_INVALID) return false;                                       // :304-309
    else if (bs.pdf < 1e-8f) return false;
    ray.delta = bs.kind == GI_BS_SPECULAR;
    if (depth > 1) throughput = throughput * (bs.bsdf / bs.pdf * (ray.delta ? 1.f : absDot(v.nrm, bs.dir)));
    ray.dir = bs.dir; ray.pdf = bs.pdf;                                               // depth 1: primSamplePdf / primSampleDelta / xv / nv, by the caller
    return true;
}
// 1: ray sampled, 0: the path ends here, -1: undecided
template <bool EXACT>
RS_D int giHead(const DevScene& s, int depth, GIVertex& v, f3& throughput, f3& Lo, Rng& rng, Stack& stack, GIRay& ray) {
    giFaceForward(v);
    if (v.mat.type != 2 && depth > 1 && !giNextEvent<EXACT>(s, v, throughput, Lo, rng, stack)) return -1;
    return giSampleBounce(depth, v, throughput, rng, ray) ? 1 : 0;
}

// the part of giTail after the bounce ray's closest hit hh is known
RS_D int giTailAfterHit(const DevScene& s, int depth, f3 curPos, const GIRay& ray, f3 throughput, f3& Lo, const Hit& hh, GIVertex& v, bool& surface) {
    surface = false;
    const f3 rd = ray.dir;
    v.wo = -rd;
    if (hh.prim < 0) {                                                                // :332-343
        if (s.envTex >= 0) {
            const f3 env = envLookup(s, rd);
            const f3 radiance = env * throughput;
            const int4 info = __ldg(s.texInfo + s.envTex);
            const float envP = luminance(env) * s.sumLightPowerInv * (float)info.x * (float)info.y * .5f;   // environmentMapPdf, scene.h:358-362
            const float weight = ray.delta ? 1.f : giPowerHeuristic(ray.pdf, envP);
            Lo = Lo + radiance * weight;
        }
        return 0;
    }
    GISurf sf;
    giSurface(s, hh, sf);
    v.pos = sf.pos; v.nrm = sf.nrm; v.mat = sf.m; v.ior = sf.ior;
    if (v.mat.type == 4) {                                                            // :346-371
        if (dot(v.nrm, rd) < 0.f) return 0;                                           // SCENE_LIGHT_SINGLE_SIDED
        const f3 radiance = v.mat.baseColor;
        float weight = 1.f;
        if (!(ray.delta || depth == 1)) {
            const float area = triangleArea(sf.v0, sf.v1, sf.v2);                     // getPrimitiveArea, scene.h:121-126
            const f3 yx = curPos - v.pos;                                             // pdfAreaToSolidAngle, mathUtil.h:182-185
            const float lp = luminance(radiance) * s.sumLightPowerInv * area * dot(yx, yx) / absDot(v.nrm, normalize(yx));
            weight = giPowerHeuristic(ray.pdf, lp);
        }
        Lo = Lo + radiance * throughput * weight;
        surface = true;
        return 0;
    }
    surface = true;
    return 1;
}

// 1: the path goes on from v, 0: it ends here, -1: undecided closest hit (never with EXACT).  surface: v.pos / v.nrm are a hit the
// sample records as xs / ns when depth == 1 (:366, :374)
template <bool EXACT>
RS_D int giTail(const DevScene& s, const GIDev& g, int depth, f3 curPos, const GIRay& ray, f3 throughput, f3& Lo, Stack& stack, const TieStore& ts,
                GIVertex& v, bool& surface) {
    const f3 ro = curPos + ray.dir * 1e-5f, rd = ray.dir;                             // makeOffsetedRay, intersections.h:12
    Hit hh;
    if (!giClosest<EXACT>(s, g, ro, rd, stack, ts, hh)) return -1;
    return giTailAfterHit(s, depth, curPos, ray, throughput, Lo, hh, v, surface);
}

// what WriteSample needs of the jittered primary hit (the reference's primMaterial / primWo / primSamplePdf / primSampleDelta)
struct GIPrim { Surf mat; f3 wo; float pdf; bool delta, shaded; };
RS_D GIPrim giEmptyPrim() {
    GIPrim p;
    p.mat.type = 0; p.mat.baseColor = mk3(0.f); p.mat.metallic = 0.f; p.mat.roughness = 0.f;
    p.wo = mk3(0.f); p.pdf = 1.f; p.delta = false; p.shaded = false;
    return p;
}

// WriteSample (:380-415)
RS_D void giWriteSample(const FrameDev& f, const GIDev& g, size_t li, const GISample& smp, const GIPrim& prim, Rng rng) {
    GIResv R;
    R.s = giEmptySample(); R.M = 0; R.w = 0.f;
    float sampleWeight = 0.f;
    if (!(luminance(smp.Lo) < 1e-8f)) {                                               // !IndirectLiSample::invalid()
        sampleWeight = luminance(smp.Lo / prim.pdf);
        if (isnan(sampleWeight) || sampleWeight < 0.f) sampleWeight = 0.f;
    }
    {
        const float r = rng.next();                                                   // Reservoir::update, restir.h:38-45
        R.w += sampleWeight; R.M++;
        if (r * R.w < sampleWeight) R.s = smp;
    }
    if (!g.first && (g.reuse & 1)) {
        const long long lli = giFindTemporal(f, li);
        GIResv T;
        T.s = giEmptySample(); T.M = 0; T.w = 0.f;
        if (lli >= 0) T = giLoad(g.resvIn, g.nszIn, (size_t)lli);
        if (!giResvInvalid(T)) {                                                      // Reservoir::merge, restir.h:61-70
            const float r = rng.next();
            R.w += T.w; R.M += T.M;
            if (r * R.w < T.w) R.s = T.s;
        }
    }
    f3 indirect = mk3(0.f);
    const GISample sample = R.s;
    if (R.M > 20) { R.w *= 20.f / R.M; R.M = 20; }                                    // clamp<20>, restir.h:88-93
    if (prim.shaded && !giResvInvalid(R)) {
        const f3 primWi = normalize(sample.xs - sample.xv);
        indirect = R.s.Lo / luminance(R.s.Lo) * R.w / (float)R.M;
        indirect = indirect * (giBSDF(prim.mat, sample.nv, prim.wo, primWi) * (prim.delta ? 1.f : satDot(sample.nv, primWi)));
    }
    if (hasNanOrInf(indirect)) indirect = mk3(0.f);
    giStore(g.resvOut, g.nszOut, li, R);
    float* out = g.indirect + 3 * li;
    const f3 prev = mk3(out[0], out[1], out[2]);
    const f3 v = (prev * (float)g.iter + indirect) / (float)(g.iter + 1);             // :415
    out[0] = v.x; out[1] = v.y; out[2] = v.z;
}

// the whole pixel after the jittered primary ray's hit is known.  false = undecided, nothing written.
template <bool EXACT>
RS_D bool giAfterHit(const DevScene& s, const FrameDev& f, const GIDev& g, int x, int y, Stack& stack, const TieStore& ts, Rng rng, f3 d, Hit h) {
    const size_t li = planeIndex(f, x, y);
    GISample smp = giEmptySample();
    GIPrim prim = giEmptyPrim();
    if (h.prim >= 0) {
        GISurf sf;
        giSurface(s, h, sf);
        if (sf.m.type != 4) {                                                         // :272
            prim.shaded = true;
            f3 throughput = mk3(1.f);
            GIVertex v;
            v.pos = sf.pos; v.nrm = sf.nrm; v.wo = -d; v.mat = sf.m; v.ior = sf.ior;
            prim.wo = v.wo;
            prim.mat = v.mat;
            for (int depth = 1; depth <= g.maxDepth; depth++) {
                GIRay ray;
                int r = giHead<EXACT>(s, depth, v, throughput, smp.Lo, rng, stack, ray);
                if (r < 0) return false;
                if (r == 0) break;
                if (depth == 1) { prim.pdf = ray.pdf; prim.delta = ray.delta; smp.xv = v.pos; smp.nv = v.nrm; }
                bool surface;
                r = giTail<EXACT>(s, g, depth, v.pos, ray, throughput, smp.Lo, stack, ts, v, surface);
                if (r < 0) return false;
                if (depth == 1 && surface) { smp.xs = v.pos; smp.ns = v.nrm; }
                if (r == 0) break;
            }
        }
    }
    giWriteSample(f, g, li, smp, prim, rng);
    return true;
}

// ------------------------------------------------------------------------------------------------ the staged form
// The same pixel as a wavefront: the vertices of the paths that are still alive are appended to a queue and the next launch runs one
// thread per live path, so the lanes of finished paths (most bounces off the ground leave the scene) do not idle through the walks of
// the others.  A launch is one iteration of the path loop -- next-event estimation at the vertex, the bounce direction, the bounce ray,
// the surface it hits -- so both walks of an iteration run with every lane of the warp alive.  Every pixel draws its random numbers in
// the order of the one-kernel form, so the two forms are bit-identical.
//   k_gi_primary   packet walk of the jittered primary rays, the surface at the hit (no per-lane walk)
//   k_gi_bounce    once per depth over the live paths: giHead + giTail
//   k_gi_resolve   WriteSample for every pixel, coherent: reservoir update, temporal merge, shading, running mean
// Between the stages, per pixel (g.pix, planes of float4): 0 {primWo, roughness} 1 {baseColor, metallic} 2 {xv, primSamplePdf}
// 3 {nv, type | delta << 8} 4 {Lo, RNG state} of the finished path 5 {xs, ns.x} 6 {ns.y, ns.z}; per live path 96 B:
// {pos, pixel} {nrm, RNG state} {wo, ior} {baseColor, metallic} {throughput, roughness} {Lo, type}.
#define RS_GI_PATH_F4 6
struct GIPathRec { float4 a, b, c, d, e, f; };
RS_D GIPathRec giPackPath(int pixel, const GIVertex& v, f3 throughput, f3 Lo, Rng rng) {
    GIPathRec r;
    r.a = make_float4(v.pos.x, v.pos.y, v.pos.z, __int_as_float(pixel));
    r.b = make_float4(v.nrm.x, v.nrm.y, v.nrm.z, __uint_as_float(rng.x));
    r.c = make_float4(v.wo.x, v.wo.y, v.wo.z, v.ior);
    r.d = make_float4(v.mat.baseColor.x, v.mat.baseColor.y, v.mat.baseColor.z, v.mat.metallic);
    r.e = make_float4(throughput.x, throughput.y, throughput.z, v.mat.roughness);
    r.f = make_float4(Lo.x, Lo.y, Lo.z, __int_as_float(v.mat.type));
    return r;
}
RS_D GIPathRec giEmptyPath() {
    GIPathRec r;
    r.a = r.b = r.c = r.d = r.e = r.f = make_float4(0.f, 0.f, 0.f, 0.f);
    return r;
}
RS_D GIPathRec giLoadPath(const float4* q, size_t i) {
    const float4* p = q + RS_GI_PATH_F4 * i;
    GIPathRec r;
    r.a = p[0]; r.b = p[1]; r.c = p[2]; r.d = p[3]; r.e = p[4]; r.f = p[5];
    return r;
}
RS_D void giPathFinished(const GIDev& g, size_t li, f3 Lo, Rng rng) { g.pix[4 * g.pixStride + li] = make_float4(Lo.x, Lo.y, Lo.z, __uint_as_float(rng.x)); }

// after the jittered primary ray's hit is known.  true: the path's first vertex leaves for bounce 1 (out)
RS_D bool giStagePrimary(const DevScene& s, const FrameDev& f, const GIDev& g, int x, int y, Rng rng, f3 d, const Hit& h, GIPathRec& out) {
    const size_t li = planeIndex(f, x, y), n = g.pixStride;
    bool shaded = false, live = false;
    if (h.prim >= 0) {
        GISurf sf;
        giSurface(s, h, sf);
        if (sf.m.type != 4) {                                                         // :272
            shaded = true;
            GIVertex v;
            v.pos = sf.pos; v.nrm = sf.nrm; v.wo = -d; v.mat = sf.m; v.ior = sf.ior;
            float4* px = g.pix + li;
            px[0] = make_float4(v.wo.x, v.wo.y, v.wo.z, v.mat.roughness);              // primWo, primMaterial
            px[n] = make_float4(v.mat.baseColor.x, v.mat.baseColor.y, v.mat.baseColor.z, v.mat.metallic);
            px[2 * n] = make_float4(0.f, 0.f, 0.f, 1.f);                                // xv, primSamplePdf: until bounce 1 samples a direction
            px[3 * n] = make_float4(0.f, 0.f, 0.f, __int_as_float(v.mat.type));        // nv, primSampleDelta
            px[5 * n] = make_float4(0.f, 0.f, 0.f, 0.f);                                // xs / ns: until bounce 1 finds a surface
            px[6 * n] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (g.maxDepth >= 1) {
                live = true;
                out = giPackPath(y * f.W + x, v, mk3(1.f), mk3(0.f), rng);
            }
        }
    }
    g.pixStatus[li] = shaded ? 1 : 0;
    if (!live) giPathFinished(g, li, mk3(0.f), rng);
    return live;
}

// iteration `depth` of the path loop for one live path.  1: the path goes on (out), 0: it ended, -1: undecided ray (pixel marked;
// the caller queues it)
RS_D int giStageBounce(const DevScene& s, const FrameDev& f, const GIDev& g, int depth, const GIPathRec& in, Stack& stack, const TieStore& ts, GIPathRec& out,
                       int& x, int& y) {
    const int pixel = __float_as_int(in.a.w);
    x = pixel % f.W; y = pixel / f.W;
    const size_t li = planeIndex(f, x, y), n = g.pixStride;
    GIVertex v;
    v.pos = mk3(in.a.x, in.a.y, in.a.z); v.nrm = mk3(in.b.x, in.b.y, in.b.z); v.wo = mk3(in.c.x, in.c.y, in.c.z); v.ior = in.c.w;
    v.mat.baseColor = mk3(in.d.x, in.d.y, in.d.z); v.mat.metallic = in.d.w; v.mat.roughness = in.e.w; v.mat.type = __float_as_int(in.f.w);
    Rng rng;
    rng.x = __float_as_uint(in.b.w);
    f3 throughput = mk3(in.e.x, in.e.y, in.e.z), Lo = mk3(in.f.x, in.f.y, in.f.z);
    GIRay ray;
    int r = giHead<false>(s, depth, v, throughput, Lo, rng, stack, ray);
    if (r < 0) { g.pixStatus[li] = 2; return -1; }
    if (r == 1) {
        if (depth == 1) {                                                             // :311-316
            g.pix[2 * n + li] = make_float4(v.pos.x, v.pos.y, v.pos.z, ray.pdf);
            g.pix[3 * n + li] = make_float4(v.nrm.x, v.nrm.y, v.nrm.z, __int_as_float(v.mat.type | (ray.delta ? 256 : 0)));
        }
        bool surface;
        r = giTail<false>(s, g, depth, v.pos, ray, throughput, Lo, stack, ts, v, surface);
        if (r < 0) { g.pixStatus[li] = 2; return -1; }
        if (depth == 1 && surface) {                                                  // :366, :374
            g.pix[5 * n + li] = make_float4(v.pos.x, v.pos.y, v.pos.z, v.nrm.x);
            g.pix[6 * n + li] = make_float4(v.nrm.y, v.nrm.z, 0.f, 0.f);
        }
        if (r == 1 && depth < g.maxDepth) {
            out = giPackPath(pixel, v, throughput, Lo, rng);
            return 1;
        }
    }
    giPathFinished(g, li, Lo, rng);
    return 0;
}

RS_D void giStageResolve(const FrameDev& f, const GIDev& g, int x, int y) {
    const size_t li = planeIndex(f, x, y), n = g.pixStride;
    const int status = g.pixStatus[li];
    if (status == 2) return;                                                          // k_restir_indirect_fix writes this pixel
    const float4* px = g.pix + li;
    GISample smp = giEmptySample();
    GIPrim prim = giEmptyPrim();
    const float4 t = px[4 * n];
    smp.Lo = mk3(t.x, t.y, t.z);
    Rng rng;
    rng.x = __float_as_uint(t.w);
    if (status == 1) {
        const float4 p0 = px[0], p1 = px[n], p2 = px[2 * n], p3 = px[3 * n], p5 = px[5 * n], p6 = px[6 * n];
        const int td = __float_as_int(p3.w);
        prim.shaded = true;
        prim.wo = mk3(p0.x, p0.y, p0.z); prim.pdf = p2.w;
        prim.mat.baseColor = mk3(p1.x, p1.y, p1.z); prim.mat.metallic = p1.w; prim.mat.roughness = p0.w;
        prim.mat.type = td & 255; prim.delta = (td & 256) != 0;
        smp.xv = mk3(p2.x, p2.y, p2.z); smp.nv = mk3(p3.x, p3.y, p3.z);
        smp.xs = mk3(p5.x, p5.y, p5.z); smp.ns = mk3(p5.w, p6.x, p6.y);
    }
    giWriteSample(f, g, li, smp, prim, rng);
}

// ------------------------------------------------------------------------------------------------ the ray-queue form
// k_gi_bounce's warps start full but run with 5-7 of 32 lanes: a warp lasts as long as its longest walk.  Here an iteration of the path
// loop is four launches and the walks are kernels of their own over ray lists, persistent warps whose lanes fetch the next ray when
// they finish one (and wait for each other at leaves), as k_shadow does for the direct path:
//   k_gi_head          giFaceForward, the light sample and what it would contribute, the bounce direction; the record is rewritten in
//                      place: {pos, pixel} {target, RNG} {dir, pdf} {c, flags} {throughput, -} {Lo, -}; slot numbers appended to the lists
//   k_gi_walk_shadow   any-hit walks of the next-event segments (depth > 1)            -> occ[slot]
//   k_gi_walk_closest  closest-hit walks of the bounce rays (leaf-box / near-tie rule)   -> hit[slot]
//   k_gi_tail          adds c if the segment was free, giTailAfterHit, survivors to the other queue
enum { GI_Q_ALIVE = 1, GI_Q_DELTA = 2, GI_Q_SHADOW = 4 };
RS_D int giQHead(const DevScene& s, const FrameDev& f, const GIDev& g, int depth, GIPathRec& rec) {
    const int pixel = __float_as_int(rec.a.w);
    GIVertex v;
    v.pos = mk3(rec.a.x, rec.a.y, rec.a.z); v.nrm = mk3(rec.b.x, rec.b.y, rec.b.z); v.wo = mk3(rec.c.x, rec.c.y, rec.c.z); v.ior = rec.c.w;
    v.mat.baseColor = mk3(rec.d.x, rec.d.y, rec.d.z); v.mat.metallic = rec.d.w; v.mat.roughness = rec.e.w; v.mat.type = __float_as_int(rec.f.w);
    Rng rng;
    rng.x = __float_as_uint(rec.b.w);
    f3 throughput = mk3(rec.e.x, rec.e.y, rec.e.z);
    giFaceForward(v);
    GINee nee;
    nee.c = mk3(0.f); nee.target = v.pos; nee.want = false;
    if (v.mat.type != 2 && depth > 1) nee = giNextEventSample(s, v, throughput, rng);
    GIRay ray;
    ray.dir = mk3(0.f); ray.pdf = 0.f; ray.delta = false;
    const bool alive = giSampleBounce(depth, v, throughput, rng, ray);
    if (alive && depth == 1) {                                                        // :311-316
        const size_t li = planeIndex(f, pixel % f.W, pixel / f.W), n = g.pixStride;
        g.pix[2 * n + li] = make_float4(v.pos.x, v.pos.y, v.pos.z, ray.pdf);
        g.pix[3 * n + li] = make_float4(v.nrm.x, v.nrm.y, v.nrm.z, __int_as_float(v.mat.type | (ray.delta ? 256 : 0)));
    }
    const int flags = (alive ? GI_Q_ALIVE : 0) | (ray.delta ? GI_Q_DELTA : 0) | (nee.want ? GI_Q_SHADOW : 0);
    rec.b = make_float4(nee.target.x, nee.target.y, nee.target.z, __uint_as_float(rng.x));
    rec.c = make_float4(ray.dir.x, ray.dir.y, ray.dir.z, ray.pdf);
    rec.d = make_float4(nee.c.x, nee.c.y, nee.c.z, __int_as_float(flags));
    rec.e = make_float4(throughput.x, throughput.y, throughput.z, 0.f);
    return flags;
}
// 1: the path goes on (out), 0: it ended, -1: undecided ray (pixel marked; the caller queues it)
RS_D int giQTail(const DevScene& s, const FrameDev& f, const GIDev& g, int depth, const GIPathRec& rec, int occ, float4 hit, GIPathRec& out, int& x, int& y) {
    const int pixel = __float_as_int(rec.a.w), flags = __float_as_int(rec.d.w);
    x = pixel % f.W; y = pixel / f.W;
    const size_t li = planeIndex(f, x, y), n = g.pixStride;
    Rng rng;
    rng.x = __float_as_uint(rec.b.w);
    const f3 throughput = mk3(rec.e.x, rec.e.y, rec.e.z);
    f3 Lo = mk3(rec.f.x, rec.f.y, rec.f.z);
    if (flags & GI_Q_SHADOW) {
        if (occ < 0) { g.pixStatus[li] = 2; return -1; }
        if (!occ) Lo = Lo + mk3(rec.d.x, rec.d.y, rec.d.z);
    }
    if (flags & GI_Q_ALIVE) {
        if (hit.w != 0.f) { g.pixStatus[li] = 2; return -1; }
        Hit hh;
        hh.bx = hit.x; hh.by = hit.y; hh.prim = __float_as_int(hit.z); hh.t = 0.f;
        GIRay ray;
        ray.dir = mk3(rec.c.x, rec.c.y, rec.c.z); ray.pdf = rec.c.w; ray.delta = (flags & GI_Q_DELTA) != 0;
        GIVertex v;
        bool surface;
        const int r = giTailAfterHit(s, depth, mk3(rec.a.x, rec.a.y, rec.a.z), ray, throughput, Lo, hh, v, surface);
        if (depth == 1 && surface) {                                                  // :366, :374
            g.pix[5 * n + li] = make_float4(v.pos.x, v.pos.y, v.pos.z, v.nrm.x);
            g.pix[6 * n + li] = make_float4(v.nrm.y, v.nrm.z, 0.f, 0.f);
        }
        if (r == 1 && depth < g.maxDepth) {
            out = giPackPath(pixel, v, throughput, Lo, rng);
            return 1;
        }
    }
    giPathFinished(g, li, Lo, rng);
    return 0;
}
// what the walkers do for one ray, in the per-lane form (tests/emu; the kernels below walk the same rays with lane refill)
RS_D float4 giQClosestOf(const DevScene& s, const GIDev& g, const GIPathRec& rec, Stack& stack, const TieStore& ts) {
    const f3 pos = mk3(rec.a.x, rec.a.y, rec.a.z), dir = mk3(rec.c.x, rec.c.y, rec.c.z);
    Hit h;
    const bool ok = giClosest<false>(s, g, pos + dir * 1e-5f, dir, stack, ts, h);
    return make_float4(h.bx, h.by, __int_as_float(h.prim), ok ? 0.f : 1.f);
}
RS_D int giQShadowOf(const DevScene& s, const GIPathRec& rec, Stack& stack) {
    return traceOccluded<false>(s, mk3(rec.a.x, rec.a.y, rec.a.z), mk3(rec.b.x, rec.b.y, rec.b.z), stack);
}

// warp-aggregated append of a slot number per wanting lane (all 32 lanes call)
RS_D void giAppendSlot(unsigned int* list, unsigned int* count, bool want, unsigned slot) {
    const unsigned m = __ballot_sync(0xffffffffu, want);
    if (!m) return;
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    unsigned base = 0;
    if (lane == leader) base = atomicAdd(count, (unsigned)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (want) list[base + __popc(m & ((1u << lane) - 1u))] = slot;
}

// warp-aggregated append of one path record per wanting lane (all 32 lanes call)
RS_D void giAppendPath(float4* q, unsigned int* count, bool want, const GIPathRec& r) {
    const unsigned m = __ballot_sync(0xffffffffu, want);
    if (!m) return;
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    unsigned base = 0;
    if (lane == leader) base = atomicAdd(count, (unsigned)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (want) {
        float4* o = q + RS_GI_PATH_F4 * (size_t)(base + __popc(m & ((1u << lane) - 1u)));
        o[0] = r.a; o[1] = r.b; o[2] = r.c; o[3] = r.d; o[4] = r.e; o[5] = r.f;
    }
}

RS_D void giPixelExact(const DevScene& s, const FrameDev& f, const CamDev& cam, const GIDev& g, int looper, int x, int y, Stack& stack) {
    Rng rng;
    f3 o, d;
    jitteredRay(f, cam, looper, x, y, rng, o, d);
    const RayT ray = makeRayT(o, d);
    Hit h;
    traceClosestExact(s, ray, h, stack);
    TieStore none;
    none.base = nullptr;
    giAfterHit<true>(s, f, g, x, y, stack, none, rng, d, h);
}

__global__ void __launch_bounds__(RS_BLOCK) k_restir_indirect(const __grid_constant__ DevScene s, const __grid_constant__ FrameDev f,
                                                              const __grid_constant__ CamDev cam, const __grid_constant__ GIDev g, int looper) {
    RS_DECLARE_STACK(stack);
    RS_DECLARE_PACKET(pk, 1);
    int x, y;
    const bool active = pixelOf(f, x, y);
    Rng rng;
    rng.x = 1;
    f3 o = mk3(0.f), d = mk3(0.f, 0.f, 1.f);
    if (active) jitteredRay(f, cam, looper, x, y, rng, o, d);
    const f3 oc = cameraOrigin(cam);
    PRay a = prayBegin(oc, d, active, pk_ta);
    packetWalk<false>(s, oc, a, a, pk_ta, pk_ta, pk_wst);
    if (active) {
        Hit h;
        if (!prayResolve(s, oc, d, a, pk_ta, h) || !giAfterHit<false>(s, f, g, x, y, stack, pk_ta, rng, d, h)) enqueuePixel(f, x, y);
    }
}
__global__ void __launch_bounds__(RS_BLOCK) k_gi_primary(const __grid_constant__ DevScene s, const __grid_constant__ FrameDev f,
                                                         const __grid_constant__ CamDev cam, const __grid_constant__ GIDev g, int looper) {
    RS_DECLARE_PACKET(pk, 1);
    int x, y;
    const bool active = pixelOf(f, x, y);
    Rng rng;
    rng.x = 1;
    f3 o = mk3(0.f), d = mk3(0.f, 0.f, 1.f);
    if (active) jitteredRay(f, cam, looper, x, y, rng, o, d);
    const f3 oc = cameraOrigin(cam);
    PRay a = prayBegin(oc, d, active, pk_ta);
    packetWalk<false>(s, oc, a, a, pk_ta, pk_ta, pk_wst);
    bool live = false;
    GIPathRec rec = giEmptyPath();
    if (active) {
        Hit h;
        if (prayResolve(s, oc, d, a, pk_ta, h)) live = giStagePrimary(s, f, g, x, y, rng, d, h, rec);
        else { g.pixStatus[planeIndex(f, x, y)] = 2; enqueuePixel(f, x, y); }
    }
    giAppendPath(g.pathQ[0], g.pathCount + 1, live, rec);
}
#ifndef RS_MINB_GI_BOUNCE
#define RS_MINB_GI_BOUNCE 6     /* 80 registers; A/B on B200 (profiles/r02_c34_gi_bench*.jsonl): 5.01 / 3.91 ms vs 5.45 / 4.34 ms uncapped (96), 5.05 / 3.84 at 8 (64) */
#endif
__global__ void __launch_bounds__(RS_BLOCK, RS_MINB_GI_BOUNCE) k_gi_bounce(const __grid_constant__ DevScene s, const __grid_constant__ FrameDev f,
                                                        const __grid_constant__ GIDev g, int depth) {
    RS_DECLARE_STACK(stack);
    RS_DECLARE_PACKET(pk, 1);
    (void)pk_tb; (void)pk_wst;
    const unsigned n = g.pathCount[depth];
    if (blockIdx.x * RS_BLOCK >= n) return;                                           // whole block out of work (uniform)
    const unsigned i = blockIdx.x * RS_BLOCK + threadIdx.x;
    bool live = false;
    GIPathRec rec = giEmptyPath();
    if (i < n) {
        const GIPathRec in = giLoadPath(g.pathQ[(depth - 1) & 1], i);
        int x, y;
        const int r = giStageBounce(s, f, g, depth, in, stack, pk_ta, rec, x, y);
        if (r < 0) enqueuePixel(f, x, y);
        live = r == 1;
    }
    giAppendPath(g.pathQ[depth & 1], g.pathCount + depth + 1, live, rec);
}
// ---- the ray-queue form's kernels
__global__ void __launch_bounds__(RS_BLOCK) k_gi_head(const __grid_constant__ DevScene s, const __grid_constant__ FrameDev f,
                                                      const __grid_constant__ GIDev g, int depth) {
    const unsigned n = g.pathCount[depth];
    if (blockIdx.x * RS_BLOCK >= n) return;                                           // whole block out of work (uniform)
    const unsigned i = blockIdx.x * RS_BLOCK + threadIdx.x;
    int flags = 0;
    if (i < n) {
        float4* q = g.pathQ[(depth - 1) & 1];
        GIPathRec rec = giLoadPath(q, i);
        flags = giQHead(s, f, g, depth, rec);
        float4* o = q + RS_GI_PATH_F4 * (size_t)i;                                    // {pos, pixel} and {Lo, -} stay
        o[1] = rec.b; o[2] = rec.c; o[3] = rec.d; o[4] = rec.e;
    }
    giAppendSlot(g.closestList, g.walkCount + 4 * depth, (flags & GI_Q_ALIVE) != 0, i);
    giAppendSlot(g.shadowList, g.walkCount + 4 * depth + 2, (flags & GI_Q_SHADOW) != 0, i);
}

#ifndef RS_MINB_GI_WALK
#define RS_MINB_GI_WALK 6
#endif
// Closest hits of the listed bounce rays: traceClosestFast's walk as a state machine, one step per round.  A lane that has finished its
// ray fetches the next one from the list as soon as RS_REFILL_MIN lanes are idle; lanes that have reached a leaf wait until RS_LEAF_MIN
// of them are there (or nobody is left to step), so the triangle tests run with several lanes.  What a ray reports does not depend on
// the order of its steps (leaf-box / near-tie rule), so the hits are traceClosestFast's.
__global__ void __launch_bounds__(RS_BLOCK, RS_MINB_GI_WALK) k_gi_walk_closest(const __grid_constant__ DevScene s, const __grid_constant__ GIDev g, int depth) {
    RS_DECLARE_STACK(stack);
    RS_DECLARE_PACKET(pk, 1);
    (void)pk_tb; (void)pk_wst;
    const unsigned FULL = 0xffffffffu;
    const unsigned n = g.walkCount[4 * depth];
    unsigned int* cursor = g.walkCount + 4 * depth + 1;
    const float4* q = g.pathQ[(depth - 1) & 1];
    const int lane = threadIdx.x & 31;
    PRay p = prayBegin(mk3(0.f), mk3(0.f, 0.f, 1.f), false, pk_ta);
    f3 o = mk3(0.f);
    int cur = RS_DONE, sp = 0;
    unsigned slot = 0;
    bool exhausted = false;                       // warp-uniform: the list is empty
    for (;;) {
        const unsigned idle = __ballot_sync(FULL, cur == RS_DONE);
        if (idle == FULL && exhausted) break;
        if (!exhausted && (__popc(idle) >= RS_REFILL_MIN || idle == FULL)) {
            unsigned base = 0;
            const int leader = __ffs(idle) - 1;
            if (lane == leader) base = atomicAdd(cursor, (unsigned)__popc(idle));
            base = __shfl_sync(FULL, base, leader);
            if (base >= n) exhausted = true;
            const unsigned i = base + __popc(idle & ((1u << lane) - 1u));
            if (cur == RS_DONE && i < n) {
                slot = g.closestList[i];
                const float4 a = q[RS_GI_PATH_F4 * (size_t)slot], c = q[RS_GI_PATH_F4 * (size_t)slot + 2];
                const f3 dir = mk3(c.x, c.y, c.z);
                o = mk3(a.x, a.y, a.z) + dir * 1e-5f;                                 // makeOffsetedRay, intersections.h:12
                bool miss = isnan(dir.x) || isnan(dir.y) || isnan(dir.z) || isnan(o.x) || isnan(o.y) || isnan(o.z);   // see traceClosestFast
                if (!miss) {
                    p = prayBegin(o, dir, true, pk_ta);
                    float t0;
                    if (slabHitP(p, s.fastRootMin[0], s.fastRootMin[1], s.fastRootMin[2], s.fastRootMax[0], s.fastRootMax[1], s.fastRootMax[2], t0)) {
                        cur = s.fastRoot; sp = 0;
                    } else miss = true;
                }
                if (miss) g.hit[slot] = make_float4(0.f, 0.f, __int_as_float(-1), 0.f);
            }
        }
        const unsigned atLeaf = __ballot_sync(FULL, cur < 0);
        const unsigned atNode = __ballot_sync(FULL, cur >= 0 && cur != RS_DONE);
        bool pop = false;
        if (atNode && __popc(atLeaf) < RS_LEAF_MIN) {
            if (cur >= 0 && cur != RS_DONE) {
                const float4* np = s.fastNodes + 4 * (size_t)cur;
                const F8 nA = ldg256(np), nB = ldg256(np + 2);
                const float4 a = nA.lo, b = nA.hi, c = nB.lo;
                const int2 l = make_int2(__float_as_int(nB.hi.x), __float_as_int(nB.hi.y));
                float tL, tR;
                const bool hL = slabHitP(p, a.x, a.y, a.z, a.w, b.x, b.y, tL);
                const bool hR = slabHitP(p, b.z, b.w, c.x, c.y, c.z, c.w, tR);
                if (hL && hR) {
                    const bool leftNear = tL <= tR;
                    stack.push(sp, leftNear ? l.y : l.x, leftNear ? tR : tL); sp++;
                    cur = leftNear ? l.x : l.y;
                } else if (hL) cur = l.x;
                else if (hR) cur = l.y;
                else pop = true;
            }
        } else if (cur < 0) {
            const int first = cur & 0x07ffffff, count = ((cur >> 27) & 7) + 1;
            const f3 d = pk_ta.dir();
            for (int k = 0; k < count; k++) {
                const Tri t = loadTriFast(s, first + k);
                prayOffer(p, o, d, pk_ta, t, first + k);
            }
            pop = true;
        }
        if (pop) {
            cur = RS_DONE;
            while (sp > 0) {
                --sp;
                if (stack.t(sp) <= p.limit) { cur = stack.ref(sp); break; }
            }
            if (cur == RS_DONE) {                                                     // the ray is finished
                Hit h;
                const bool ok = prayResolve(s, o, pk_ta.dir(), p, pk_ta, h);
                g.hit[slot] = make_float4(h.bx, h.by, __int_as_float(h.prim), ok ? 0.f : 1.f);
            }
        }
    }
}

// The listed next-event segments, any hit: k_shadow's loop (refill, leaf wait, nearer child first) on traceOccluded's ray
__global__ void __launch_bounds__(RS_BLOCK, RS_MINB_SHADOW) k_gi_walk_shadow(const __grid_constant__ DevScene s, const __grid_constant__ GIDev g, int depth) {
    RS_DECLARE_REFSTACK(stack);
    const unsigned FULL = 0xffffffffu;
    const unsigned n = g.walkCount[4 * depth + 2];
    unsigned int* cursor = g.walkCount + 4 * depth + 3;
    const float4* q = g.pathQ[(depth - 1) & 1];
    const int lane = threadIdx.x & 31;
    RayT r;
    RayF rf;
    float dist = 0.f;
    int cur = RS_DONE, sp = 0;
    unsigned slot = 0;
    bool exhausted = false;
    r.o = r.d = r.inv = mk3(0.f); r.flags = 0; r.dim = r.lesser = 0;
    rf.inv = rf.oi = mk3(0.f);
    for (;;) {
        const unsigned idle = __ballot_sync(FULL, cur == RS_DONE);
        if (idle == FULL && exhausted) break;
        if (!exhausted && (__popc(idle) >= RS_REFILL_MIN || idle == FULL)) {
            unsigned base = 0;
            const int leader = __ffs(idle) - 1;
            if (lane == leader) base = atomicAdd(cursor, (unsigned)__popc(idle));
            base = __shfl_sync(FULL, base, leader);
            if (base >= n) exhausted = true;
            const unsigned i = base + __popc(idle & ((1u << lane) - 1u));
            if (cur == RS_DONE && i < n) {
                slot = g.shadowList[i];
                const float4 a = q[RS_GI_PATH_F4 * (size_t)slot], b = q[RS_GI_PATH_F4 * (size_t)slot + 1];
                const f3 x = mk3(a.x, a.y, a.z);
                f3 dir = mk3(b.x, b.y, b.z) - x;                                      // traceOccluded (scene.h:286-295)
                float dd = length(dir);
                if (dd > 0.f) {
                    dir = dir / dd;
                    r = makeRayT(x + dir * 1e-5f, dir);
                    dist = dd - 1e-4f * 2.f;
                    rf = makeRayF(r);
                    float tr;
                    if (slabHit(rf, s.fastRootMin[0], s.fastRootMin[1], s.fastRootMin[2], s.fastRootMax[0], s.fastRootMax[1], s.fastRootMax[2], dist, tr)) {
                        cur = s.fastRoot; sp = 0;
                    }
                }
                if (cur == RS_DONE) g.occ[slot] = 0;
            }
        }
        const unsigned atLeaf = __ballot_sync(FULL, cur < 0);
        const unsigned atNode = __ballot_sync(FULL, cur >= 0 && cur != RS_DONE);
        if (atNode && __popc(atLeaf) < RS_LEAF_MIN) {
            if (cur >= 0 && cur != RS_DONE) {
                const float4* np = s.fastNodes + 4 * (size_t)cur;
                const F8 nA = ldg256(np), nB = ldg256(np + 2);
                const float4 a = nA.lo, b = nA.hi, c = nB.lo;
                const int2 l = make_int2(__float_as_int(nB.hi.x), __float_as_int(nB.hi.y));
                float tL, tR;
                const bool hL = slabHit(rf, a.x, a.y, a.z, a.w, b.x, b.y, dist, tL);
                const bool hR = slabHit(rf, b.z, b.w, c.x, c.y, c.z, c.w, dist, tR);
                if (hL && hR) { const bool ln = tL <= tR; stack.pushRef(sp, ln ? l.y : l.x); sp++; cur = ln ? l.x : l.y; }
                else if (hL) cur = l.x;
                else if (hR) cur = l.y;
                else cur = sp == 0 ? RS_DONE : stack.ref(--sp);
                if (cur == RS_DONE) g.occ[slot] = 0;
            }
        } else if (cur < 0) {
            if (leafOccluded(s, r, dist, cur)) { g.occ[slot] = 1; cur = RS_DONE; }
            else {
                cur = sp == 0 ? RS_DONE : stack.ref(--sp);
                if (cur == RS_DONE) g.occ[slot] = 0;
            }
        }
    }
}

__global__ void __launch_bounds__(RS_BLOCK) k_gi_tail(const __grid_constant__ DevScene s, const __grid_constant__ FrameDev f,
                                                      const __grid_constant__ GIDev g, int depth) {
    const unsigned n = g.pathCount[depth];
    if (blockIdx.x * RS_BLOCK >= n) return;
    const unsigned i = blockIdx.x * RS_BLOCK + threadIdx.x;
    bool live = false;
    GIPathRec rec = giEmptyPath();
    if (i < n) {
        const GIPathRec in = giLoadPath(g.pathQ[(depth - 1) & 1], i);
        const int flags = __float_as_int(in.d.w);
        const int occ = (flags & GI_Q_SHADOW) ? g.occ[i] : 0;
        const float4 hit = (flags & GI_Q_ALIVE) ? g.hit[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        int x, y;
        const int r = giQTail(s, f, g, depth, in, occ, hit, rec, x, y);
        if (r < 0) enqueuePixel(f, x, y);
        live = r == 1;
    }
    giAppendPath(g.pathQ[depth & 1], g.pathCount + depth + 1, live, rec);
}

__global__ void __launch_bounds__(RS_BLOCK) k_gi_resolve(const __grid_constant__ FrameDev f, const __grid_constant__ GIDev g) {
    int x, y;
    if (pixelOf(f, x, y)) giStageResolve(f, g, x, y);
}
__global__ void __launch_bounds__(RS_BLOCK) k_restir_indirect_exact(const __grid_constant__ DevScene s, const __grid_constant__ FrameDev f,
                                                                    const __grid_constant__ CamDev cam, const __grid_constant__ GIDev g, int looper) {
    RS_DECLARE_STACK(stack);
    int x, y;
    if (pixelOf(f, x, y)) giPixelExact(s, f, cam, g, looper, x, y, stack);
}
__global__ void __launch_bounds__(RS_BLOCK) k_restir_indirect_fix(const __grid_constant__ DevScene s, const __grid_constant__ FrameDev f,
                                                                  const __grid_constant__ CamDev cam, const __grid_constant__ GIDev g, int looper) {
    RS_DECLARE_STACK(stack);
    const unsigned n = *f.queueCount;
    const unsigned numWarps = gridDim.x * (RS_BLOCK / 32), gw = blockIdx.x * (RS_BLOCK / 32) + (threadIdx.x >> 5);
    for (unsigned i = (threadIdx.x & 31) * numWarps + gw; i < n; i += 32 * numWarps) {
        const int idx = f.queue[i];
        giPixelExact(s, f, cam, g, looper, idx % f.W, idx / f.W, stack);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && g.fallback) atomicAdd(g.fallback, n);
}

// device reservoirs -> the reference's Reservoir<IndirectLiSample> (17 x 4 bytes: Lo xv nv xs ns, numSamples (int), weight)
__global__ void k_export_gi(const float4* __restrict__ rec, const float* __restrict__ nsz, float* out17, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const GIResv r = giLoad(rec, nsz, i);
    float* o = out17 + 17 * i;
    o[0] = r.s.Lo.x; o[1] = r.s.Lo.y; o[2] = r.s.Lo.z; o[3] = r.s.xv.x; o[4] = r.s.xv.y; o[5] = r.s.xv.z;
    o[6] = r.s.nv.x; o[7] = r.s.nv.y; o[8] = r.s.nv.z; o[9] = r.s.xs.x; o[10] = r.s.xs.y; o[11] = r.s.xs.z;
    o[12] = r.s.ns.x; o[13] = r.s.ns.y; o[14] = r.s.ns.z;
    o[15] = __int_as_float(r.M); o[16] = r.w;
}

#ifndef RS_HOST_EMU
int launchRestirIndirect(const DevScene& s, const FrameDev& f, const CamDev& cam, const GIDev& g, int looper, int numSMs, cudaStream_t st) {
    if (s.traversal == RS_TRAVERSAL_EXACT) { k_restir_indirect_exact<<<pixelGrid(f), RS_BLOCK, 0, st>>>(s, f, cam, g, looper); return 1; }
    cudaMemsetAsync(f.queueCount, 0, sizeof(unsigned int), st);
    if (g.pix) {                                                                      // staged: one launch per bounce over the live paths
        cudaMemsetAsync(g.pathCount, 0, (size_t)(g.maxDepth + 2) * sizeof(unsigned int), st);
        k_gi_primary<<<pixelGrid(f), RS_BLOCK, 0, st>>>(s, f, cam, g, looper);
        const unsigned blocks = (unsigned)((g.pixStride + RS_BLOCK - 1) / RS_BLOCK);   // the path count lives on the device: blocks past it return at once
        int launches = 3;
        if (g.hit) {                                                                  // ray queues: the walks as persistent kernels with lane refill
            cudaMemsetAsync(g.walkCount, 0, (size_t)(4 * (g.maxDepth + 1)) * sizeof(unsigned int), st);
            for (int depth = 1; depth <= g.maxDepth; depth++) {
                k_gi_head<<<blocks, RS_BLOCK, 0, st>>>(s, f, g, depth);
                if (depth > 1) k_gi_walk_shadow<<<(unsigned)(numSMs * RS_MINB_SHADOW), RS_BLOCK, 0, st>>>(s, g, depth);
                k_gi_walk_closest<<<(unsigned)(numSMs * RS_MINB_GI_WALK), RS_BLOCK, 0, st>>>(s, g, depth);
                k_gi_tail<<<blocks, RS_BLOCK, 0, st>>>(s, f, g, depth);
                launches += depth > 1 ? 4 : 3;
            }
        } else {
            for (int depth = 1; depth <= g.maxDepth; depth++) k_gi_bounce<<<blocks, RS_BLOCK, 0, st>>>(s, f, g, depth);
            launches += g.maxDepth;
        }
        k_gi_resolve<<<pixelGrid(f), RS_BLOCK, 0, st>>>(f, g);
        k_restir_indirect_fix<<<RS_FIX_BLOCKS, RS_BLOCK, 0, st>>>(s, f, cam, g, looper);
        return launches;
    }
    k_restir_indirect<<<pixelGrid(f), RS_BLOCK, 0, st>>>(s, f, cam, g, looper);
    k_restir_indirect_fix<<<RS_FIX_BLOCKS, RS_BLOCK, 0, st>>>(s, f, cam, g, looper);
    return 2;
}
void launchExportGI(const float4* rec, const float* nsz, float* out17, size_t n, cudaStream_t st) {
    k_export_gi<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(rec, nsz, out17, n);
}
#endif  // RS_HOST_EMU
