// oracle_capi.cpp -- C entry points over lyon_oracle.hpp for ctypes (tests, smoke, bench cpu_baseline).
// TEST INFRASTRUCTURE ONLY: nothing under lyon_b200/ may link or load this library.
//
// Batch semantics follow SURVEY §8(b): a batch is `for p in paths { StrokeTessellator::new()
// .tessellate_path(p, opts, &mut BuffersBuilder::new(&mut out, ctor)) }` on ONE VertexBuffers, so
// vertex ids are absolute (geometry_builder.rs:342-349).
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <thread>

#include "lyon_oracle.hpp"

using namespace lyon_oracle;

struct oracle_mesh {
    Buffers buf;
    std::vector<uint64_t> path_vertex_end;  // per path: vertices.size() after the path
    std::vector<uint64_t> path_index_end;
};

extern "C" {

// Stroke a batch on one thread into one buffer. Returns 0 or a negative error (first failing path aborts the batch).
// max_index: largest vertex id that may be handed out (0xFFFFFFFE for u32 buffers).
oracle_mesh* oracle_stroke_batch(const uint8_t* verbs, const float* points, const uint32_t* path_verb_begin,
                                 const uint32_t* path_point_begin, uint32_t n_paths, uint32_t num_attributes,
                                 const Options* opts, uint64_t max_index, int* err) {
    oracle_mesh* m = new oracle_mesh();
    m->buf.max_index = max_index;
    *err = OK;
    for (uint32_t p = 0; p < n_paths; ++p) {
        uint32_t vb = path_verb_begin[p], ve = path_verb_begin[p + 1];
        int e = stroke_path(verbs + vb, ve - vb, points + 2 * (size_t)path_point_begin[p], num_attributes, *opts, m->buf);
        if (e != OK) {
            *err = e;
            // tessellate_paths returns Err: nothing is handed back.
            m->buf.vertices.clear();
            m->buf.indices.clear();
            m->path_vertex_end.clear();
            m->path_index_end.clear();
            return m;
        }
        m->path_vertex_end.push_back(m->buf.vertices.size());
        m->path_index_end.push_back(m->buf.indices.size());
    }
    return m;
}
uint64_t oracle_mesh_num_vertices(const oracle_mesh* m) { return m->buf.vertices.size(); }
uint64_t oracle_mesh_num_indices(const oracle_mesh* m) { return m->buf.indices.size(); }
void oracle_mesh_copy(const oracle_mesh* m, void* vertices, uint32_t* indices, uint64_t* path_vertex_end,
                      uint64_t* path_index_end) {
    if (vertices) memcpy(vertices, m->buf.vertices.data(), m->buf.vertices.size() * sizeof(Vertex));
    if (indices) memcpy(indices, m->buf.indices.data(), m->buf.indices.size() * sizeof(uint32_t));
    if (path_vertex_end) memcpy(path_vertex_end, m->path_vertex_end.data(), m->path_vertex_end.size() * 8);
    if (path_index_end) memcpy(path_index_end, m->path_index_end.data(), m->path_index_end.size() * 8);
}
void oracle_mesh_free(oracle_mesh* m) { delete m; }

// StrokeVertex::interpolated_attributes (stroke.rs:2728-2747) of every vertex of the mesh, against the paths' own
// attribute stores (Path as AttributeStore: the f32s stored after the endpoint, path.rs:650-665):
// Endpoint{id} -> store.get(id); Edge{from, to, t} -> a[i] * (1.0 - t) + b[i] * t.  out[v * num_attributes + i].
void oracle_mesh_interpolated_attributes(const oracle_mesh* m, const float* points, const uint32_t* path_point_begin,
                                         uint32_t num_attributes, float* out) {
    size_t v = 0;
    for (size_t p = 0; p < m->path_vertex_end.size(); ++p) {
        const float* store = points + 2 * (size_t)path_point_begin[p];
        for (; v < m->path_vertex_end[p]; ++v) {
            const Vertex& vx = m->buf.vertices[v];
            const float* b = store + 2 * ((size_t)(vx.src & 0x7FFFFFFFu) + 1);
            float* o = out + v * num_attributes;
            if (vx.t == 1.0f) {
                for (uint32_t i = 0; i < num_attributes; ++i) o[i] = b[i];
            } else {
                const float* a = store + 2 * ((size_t)m->buf.vertex_from[v] + 1);
                const float t = vx.t;
                for (uint32_t i = 0; i < num_attributes; ++i) o[i] = a[i] * (1.0f - t) + b[i] * t;
            }
        }
    }
}

// Timed multi-threaded run for the CPU baseline ("rayon-parallel over paths" stand-in): the batch is cut
// into n_threads contiguous ranges of equal verb count; every worker tessellates its range into its own
// VertexBuffers.  Like bench/tess (which reuses a pre-sized VertexBuffers across iterations) the buffers are
// warm: one untimed pass sizes and touches them, they are cleared (capacity kept), and the second pass is
// timed.  Returns seconds of wall clock of the timed parallel region; writes total vertices / indices.
double oracle_stroke_batch_timed(const uint8_t* verbs, const float* points, const uint32_t* path_verb_begin,
                                 const uint32_t* path_point_begin, uint32_t n_paths, uint32_t num_attributes,
                                 const Options* opts, int n_threads, uint64_t* total_vertices,
                                 uint64_t* total_indices) {
    if (n_threads < 1) n_threads = 1;
    std::vector<Buffers> bufs((size_t)n_threads);
    std::vector<uint32_t> cut((size_t)n_threads + 1, n_paths);
    cut[0] = 0;
    {
        const uint64_t total_verbs = path_verb_begin[n_paths];
        uint32_t p = 0;
        for (int t = 1; t < n_threads; ++t) {
            uint64_t target = total_verbs * (uint64_t)t / (uint64_t)n_threads;
            while (p < n_paths && path_verb_begin[p] < target) ++p;
            cut[(size_t)t] = p;
        }
    }
    double secs = 0.0;
    for (int pass = 0; pass < 2; ++pass) {
        for (auto& b : bufs) {
            b.vertices.clear();
            b.indices.clear();
        }
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; ++t) {
            th.emplace_back([&, t]() {
                Buffers& b = bufs[(size_t)t];
                for (uint32_t p = cut[(size_t)t]; p < cut[(size_t)t + 1]; ++p) {
                    uint32_t vb = path_verb_begin[p], ve = path_verb_begin[p + 1];
                    stroke_path(verbs + vb, ve - vb, points + 2 * (size_t)path_point_begin[p], num_attributes, *opts, b);
                }
            });
        }
        for (auto& x : th) x.join();
        auto t1 = std::chrono::steady_clock::now();
        secs = std::chrono::duration<double>(t1 - t0).count();
    }
    uint64_t tv = 0, ti = 0;
    for (auto& b : bufs) {
        tv += b.vertices.size();
        ti += b.indices.size();
    }
    *total_vertices = tv;
    *total_indices = ti;
    return secs;
}

// ---- flattening -------------------------------------------------------------------------------
// Flatten a batch of standalone cubic segments (ctrl = 8 floats per curve: from, ctrl1, ctrl2, to), the
// point sequence of CubicBezierSegment::for_each_flattened (cubic_bezier.rs:563-582): for each curve the
// points AFTER `from`.  offsets_out has n+1 entries.  If points_out is NULL only offsets are written.
void oracle_flatten_cubics(const float* ctrl, uint64_t n, float tol, uint64_t* offsets_out, float* points_out,
                           float* t_out) {
    uint64_t off = 0;
    for (uint64_t i = 0; i < n; ++i) {
        const float* c = ctrl + 8 * i;
        offsets_out[i] = off;
        cubic_for_each_flattened_with_t(v2(c[0], c[1]), v2(c[2], c[3]), v2(c[4], c[5]), v2(c[6], c[7]), tol,
                                        [&](V2 p, float t) {
                                            if (points_out) {
                                                points_out[2 * off] = p.x;
                                                points_out[2 * off + 1] = p.y;
                                            }
                                            if (t_out) t_out[off] = t;
                                            off++;
                                        });
    }
    offsets_out[n] = off;
}
// Same for quadratic segments (ctrl = 6 floats per curve), QuadraticBezierSegment::for_each_flattened
// (quadratic_bezier.rs:337-360).
void oracle_flatten_quadratics(const float* ctrl, uint64_t n, float tol, uint64_t* offsets_out, float* points_out,
                               float* t_out) {
    uint64_t off = 0;
    for (uint64_t i = 0; i < n; ++i) {
        const float* c = ctrl + 6 * i;
        offsets_out[i] = off;
        quad_for_each_flattened_with_t(v2(c[0], c[1]), v2(c[2], c[3]), v2(c[4], c[5]), tol, [&](V2 p, float t) {
            if (points_out) {
                points_out[2 * off] = p.x;
                points_out[2 * off + 1] = p.y;
            }
            if (t_out) t_out[off] = t;
            off++;
        });
    }
    offsets_out[n] = off;
}
double oracle_flatten_cubics_timed(const float* ctrl, uint64_t n, float tol, int n_threads, float* points_out,
                                   const uint64_t* offsets, uint64_t* total) {
    if (n_threads < 1) n_threads = 1;
    std::atomic<uint64_t> next(0);
    std::atomic<uint64_t> tot(0);
    const uint64_t block = 4096;
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t) {
        th.emplace_back([&]() {
            uint64_t local = 0;
            for (;;) {
                uint64_t s = next.fetch_add(block);
                if (s >= n) break;
                uint64_t e = s + block < n ? s + block : n;
                for (uint64_t i = s; i < e; ++i) {
                    const float* c = ctrl + 8 * i;
                    uint64_t off = offsets[i];
                    cubic_for_each_flattened_with_t(v2(c[0], c[1]), v2(c[2], c[3]), v2(c[4], c[5]), v2(c[6], c[7]),
                                                    tol, [&](V2 p, float) {
                                                        points_out[2 * off] = p.x;
                                                        points_out[2 * off + 1] = p.y;
                                                        off++;
                                                        local++;
                                                    });
                }
            }
            tot += local;
        });
    }
    for (auto& x : th) x.join();
    auto t1 = std::chrono::steady_clock::now();
    *total = tot.load();
    return std::chrono::duration<double>(t1 - t0).count();
}

// Path-level flattening (crates/path/src/iterator.rs:150-207 `Flattened`): every curve event becomes
// the line_to sequence of its flattened points; Begin/Line/End/Close pass through.  Works on paths
// WITHOUT custom attributes.  Output verbs/points use the same encoding as the input.
// out arrays must be large enough (call once with NULL outputs to get the sizes).
// Path-level flattening with custom attributes: the `Flattened` builder adapter's helpers (crates/path/src/private.rs:56-122,
// flatten_quadratic_bezier / flatten_cubic_bezier): every flattened point of a curve carries
// prev_attributes[i] * (1 - t) + attributes[i] * t, the curve's own attributes where t == 1.0; the other verbs pass through
// with their attributes.  Attributes are stored two per point slot after every endpoint (path.rs:650-665), an odd
// count padded with 0.
void oracle_flatten_paths_attr(const uint8_t* verbs, const float* points, const uint32_t* path_verb_begin,
                               const uint32_t* path_point_begin, uint32_t n_paths, uint32_t num_attributes, float tol,
                               uint8_t* out_verbs, float* out_points, uint32_t* out_path_verb_begin,
                               uint32_t* out_path_point_begin, uint64_t* n_out_verbs, uint64_t* n_out_points) {
    uint64_t ov = 0, op = 0;
    const uint32_t A = num_attributes, stride = (A + 1u) / 2u;
    std::vector<float> buf(2 * (size_t)stride + 2, 0.0f);
    for (uint32_t p = 0; p < n_paths; ++p) {
        if (out_path_verb_begin) out_path_verb_begin[p] = (uint32_t)ov;
        if (out_path_point_begin) out_path_point_begin[p] = (uint32_t)op;
        const float* pts = points + 2 * (size_t)path_point_begin[p];
        uint32_t cur = 0;
        V2 current = v2(0, 0), first = v2(0, 0);
        const float* cur_attr = nullptr;    // attributes of the current endpoint
        const float* first_attr = nullptr;  // attributes of the sub-path's first endpoint
        auto P = [&](uint32_t i) { return v2(pts[2 * i], pts[2 * i + 1]); };
        auto emit = [&](uint8_t verb, bool has_pt, V2 pt, const float* attr) {
            if (out_verbs) out_verbs[ov] = verb;
            ov++;
            if (has_pt) {
                if (out_points) {
                    out_points[2 * op] = pt.x;
                    out_points[2 * op + 1] = pt.y;
                    for (uint32_t i = 0; i < 2 * stride; ++i) out_points[2 * (op + 1) + i] = i < A ? attr[i] : 0.0f;
                }
                op += 1 + stride;
            }
        };
        auto lerp_attr = [&](const float* prev, const float* attr, float t) -> const float* {
            if (t == 1.0f) return attr;
            for (uint32_t i = 0; i < A; ++i) buf[i] = prev[i] * (1.0f - t) + attr[i] * t;
            return buf.data();
        };
        for (uint32_t vi = path_verb_begin[p]; vi < path_verb_begin[p + 1]; ++vi) {
            switch (verbs[vi]) {
                case 3:
                    current = first = P(cur);
                    cur_attr = first_attr = pts + 2 * (size_t)(cur + 1);
                    cur += 1 + stride;
                    emit(3, true, current, cur_attr);
                    break;
                case 0:
                    current = P(cur);
                    cur_attr = pts + 2 * (size_t)(cur + 1);
                    cur += 1 + stride;
                    emit(0, true, current, cur_attr);
                    break;
                case 1: {
                    V2 c = P(cur), to = P(cur + 1);
                    const float* attr = pts + 2 * (size_t)(cur + 2);
                    const float* prev = cur_attr;
                    cur += 2 + stride;
                    quad_for_each_flattened_with_t(current, c, to, tol, [&](V2 q, float t) { emit(0, true, q, lerp_attr(prev, attr, t)); });
                    current = to;
                    cur_attr = attr;
                    break;
                }
                case 2: {
                    V2 c1 = P(cur), c2 = P(cur + 1), to = P(cur + 2);
                    const float* attr = pts + 2 * (size_t)(cur + 3);
                    const float* prev = cur_attr;
                    cur += 3 + stride;
                    cubic_for_each_flattened_with_t(current, c1, c2, to, tol, [&](V2 q, float t) { emit(0, true, q, lerp_attr(prev, attr, t)); });
                    current = to;
                    cur_attr = attr;
                    break;
                }
                case 4:
                    cur += 1 + stride;
                    emit(4, true, first, first_attr);
                    current = first;
                    cur_attr = first_attr;
                    break;
                case 5:
                    emit(5, false, first, nullptr);
                    current = first;
                    cur_attr = first_attr;
                    break;
            }
        }
    }
    if (out_path_verb_begin) out_path_verb_begin[n_paths] = (uint32_t)ov;
    if (out_path_point_begin) out_path_point_begin[n_paths] = (uint32_t)op;
    *n_out_verbs = ov;
    *n_out_points = op;
}
void oracle_flatten_paths(const uint8_t* verbs, const float* points, const uint32_t* path_verb_begin,
                          const uint32_t* path_point_begin, uint32_t n_paths, float tol, uint8_t* out_verbs,
                          float* out_points, uint32_t* out_path_verb_begin, uint32_t* out_path_point_begin,
                          uint64_t* n_out_verbs, uint64_t* n_out_points) {
    oracle_flatten_paths_attr(verbs, points, path_verb_begin, path_point_begin, n_paths, 0u, tol, out_verbs, out_points,
                              out_path_verb_begin, out_path_point_begin, n_out_verbs, n_out_points);
}

// PathMeasurements::initialize (crates/algorithms/src/measure.rs:170-257): the edges of every path of a batch.  Pass null
// outputs to count.  distance / index / t: one entry per edge; path_edge_begin: n_paths + 1; path_length: n_paths.
void oracle_measure_paths(const uint8_t* verbs, const float* points, const uint32_t* path_verb_begin,
                          const uint32_t* path_point_begin, uint32_t n_paths, uint32_t num_attributes, float tolerance,
                          float* distance_out, uint32_t* index_out, float* t_out, uint64_t* path_edge_begin, float* path_length,
                          uint64_t* n_edges) {
    const float tol = tolerance > 1e-4f ? tolerance : 1e-4f;  // tolerance.max(1e-4)
    const uint32_t stride = (num_attributes + 1u) / 2u;
    uint64_t ne = 0;
    for (uint32_t p = 0; p < n_paths; ++p) {
        if (path_edge_begin) path_edge_begin[p] = ne;
        const float* pts = points + 2 * (size_t)path_point_begin[p];
        uint32_t cur = 0, index = 0;
        V2 current = v2(0, 0), first = v2(0, 0);
        float distance = 0.0f;
        auto P = [&](uint32_t i) { return v2(pts[2 * i], pts[2 * i + 1]); };
        auto push = [&](float t) {
            if (distance_out) {
                distance_out[ne] = distance;
                index_out[ne] = index;
                t_out[ne] = t;
            }
            ne++;
        };
        for (uint32_t vi = path_verb_begin[p]; vi < path_verb_begin[p + 1]; ++vi, ++index) {
            switch (verbs[vi]) {
                case 3:  // Begin
                    current = first = P(cur);
                    cur += 1 + stride;
                    push(1.0f);
                    break;
                case 0: {  // Line: distance += (from - to).length()
                    V2 to = P(cur);
                    cur += 1 + stride;
                    distance = distance + length(current - to);
                    current = to;
                    push(1.0f);
                    break;
                }
                case 1: {
                    V2 c = P(cur), to = P(cur + 1);
                    cur += 2 + stride;
                    V2 prev = current;
                    quad_for_each_flattened_with_t(current, c, to, tol, [&](V2 q, float t) {
                        distance = distance + length(q - prev);  // line.length() = (to - from).length()
                        prev = q;
                        push(t);
                    });
                    current = to;
                    break;
                }
                case 2: {
                    V2 c1 = P(cur), c2 = P(cur + 1), to = P(cur + 2);
                    cur += 3 + stride;
                    V2 prev = current;
                    cubic_for_each_flattened_with_t(current, c1, c2, to, tol, [&](V2 q, float t) {
                        distance = distance + length(q - prev);
                        prev = q;
                        push(t);
                    });
                    current = to;
                    break;
                }
                case 4:  // End { close: true }: distance += (last - first).length()
                    cur += 1 + stride;
                    distance = distance + length(current - first);
                    current = first;
                    push(1.0f);
                    break;
                case 5:  // End { close: false }: no edge
                    current = first;
                    break;
            }
        }
        if (path_length) path_length[p] = distance;
    }
    if (path_edge_begin) path_edge_begin[n_paths] = ne;
    *n_edges = ne;
}

// ---- shape helpers of the path builders (crates/path/src/builder.rs:555-700, 1506-1660) and StrokeBuilder::add_rectangle's
// thin-rectangle rule (crates/tessellation/src/stroke.rs:523-537, 2336-2369), restated for the host-side generators of
// lyon_b200/shapes.py.  A shape is returned as lyon Path storage: verbs + points (no attributes).
namespace {
struct ShapeOut {
    uint8_t* verbs;
    float* points;
    uint32_t nv = 0, np = 0;
    void pt(V2 p) {
        if (points) {
            points[2 * np] = p.x;
            points[2 * np + 1] = p.y;
        }
        np++;
    }
    void verb(uint8_t v) {
        if (verbs) verbs[nv] = v;
        nv++;
    }
    V2 first = v2(0, 0);
    void begin(V2 p) { first = p; pt(p); verb(3); }
    void line_to(V2 p) { pt(p); verb(0); }
    void quad_to(V2 c, V2 p) { pt(c); pt(p); verb(1); }
    void cubic_to(V2 c1, V2 c2, V2 p) { pt(c1); pt(c2); pt(p); verb(2); }
    void end(bool close) {
        if (close) pt(first);
        verb(close ? 4 : 5);
    }
};
// sample_ellipse (crates/geom/src/arc.rs:745-750): Rotation::new(x_rotation).transform_point(rx cos a, ry sin a)
V2 shape_sample_ellipse(V2 radii, float x_rotation, float angle) {
    const float ex = radii.x * libm_cos(angle), ey = radii.y * libm_sin(angle);
    const float c = libm_cos(x_rotation), s = libm_sin(x_rotation);
    return v2(ex * c - ey * s, ey * c + ex * s);
}
}  // namespace

// kind 0 rectangle (p: min.x min.y max.x max.y), 1 circle (cx cy r), 2 ellipse (cx cy rx ry x_rotation),
// 3 rounded rectangle (min.x min.y max.x max.y tl tr bl br); winding +1 / -1.  Null outputs count.
void oracle_shape_path(int kind, const float* p, int winding, uint8_t* verbs, float* points, uint32_t* n_verbs, uint32_t* n_points) {
    ShapeOut o{verbs, points};
    const bool positive = winding >= 0;
    const float dir = positive ? 1.0f : -1.0f;
    if (kind == 0) {  // builder.rs:649-678
        const V2 mn = v2(p[0], p[1]), mx = v2(p[2], p[3]);
        o.begin(mn);
        if (positive) {
            o.line_to(v2(mx.x, mn.y));
            o.line_to(mx);
            o.line_to(v2(mn.x, mx.y));
        } else {
            o.line_to(v2(mn.x, mx.y));
            o.line_to(mx);
            o.line_to(v2(mx.x, mn.y));
        }
        o.end(true);
    } else if (kind == 1) {  // builder.rs:1508-1546
        const V2 c = v2(p[0], p[1]);
        const float radius = fabsf(p[2]);
        const float d = radius * 0.55191505f;
        o.begin(c + v2(-radius, 0.0f));
        o.cubic_to(c + v2(-radius, -d * dir), c + v2(-d, -radius * dir), c + v2(0.0f, -radius * dir));
        o.cubic_to(c + v2(d, -radius * dir), c + v2(radius, -d * dir), c + v2(radius, 0.0f));
        o.cubic_to(c + v2(radius, d * dir), c + v2(d, radius * dir), c + v2(0.0f, radius * dir));
        o.cubic_to(c + v2(-d, radius * dir), c + v2(-radius, d * dir), c + v2(-radius, 0.0f));
        o.end(true);
    } else if (kind == 2) {  // builder.rs:606-636 + arc.rs:668-703 (arc_to_quadratic_beziers_with_t)
        const V2 c = v2(p[0], p[1]), radii = v2(p[2], p[3]);
        const float x_rotation = p[4], start_angle = 0.0f;
        const float sweep_angle = (2.0f * PI_F) * dir;
        o.begin(c + shape_sample_ellipse(radii, x_rotation, start_angle + sweep_angle * 0.0f));
        const float sign = sweep_angle < 0.0f ? -1.0f : 1.0f;
        const float sweep = fminf(fabsf(sweep_angle), PI_F * 2.0f);
        const float n_steps = ceilf(sweep / (PI_F / 4.0f));
        const float step = sweep / n_steps * sign;
        const float alpha = (float)tan((double)(step * 0.5f));
        const int n = (int)n_steps;
        const float xc = libm_cos(x_rotation), xs = libm_sin(x_rotation);
        for (int i = 0; i < n; ++i) {
            const float a1 = start_angle + step * (float)i, a2 = start_angle + step * (float)(i + 1);
            const V2 from = c + shape_sample_ellipse(radii, x_rotation, a1);
            const V2 to = c + shape_sample_ellipse(radii, x_rotation, a2);
            // Arc::tangent_at_angle (arc.rs:231-243): rotate (-rx sin a, ry cos a) by x_rotation
            const float tx = -radii.x * libm_sin(a1), ty = radii.y * libm_cos(a1);
            const V2 tg = v2(tx * xc - ty * xs, ty * xc + tx * xs);
            o.quad_to(from + tg * alpha, to);
        }
        o.end(true);
    } else if (kind == 3) {  // builder.rs:1551-1660
        const float x_min = p[0], y_min = p[1], x_max = p[2], y_max = p[3];
        const float w = x_max - x_min, h = y_max - y_min, min_wh = fminf(w, h);
        float tl = fminf(fabsf(p[4]), min_wh), tr = fminf(fabsf(p[5]), min_wh), bl = fminf(fabsf(p[6]), min_wh),
              br = fminf(fabsf(p[7]), min_wh);
        if (tl + tr > w) { const float x = (tl + tr - w) * 0.5f; tl -= x; tr -= x; }
        if (bl + br > w) { const float x = (bl + br - w) * 0.5f; bl -= x; br -= x; }
        if (tr + br > h) { const float x = (tr + br - h) * 0.5f; tr -= x; br -= x; }
        if (tl + bl > h) { const float x = (tl + bl - h) * 0.5f; tl -= x; bl -= x; }
        const float K = 0.55191505f;
        const float tl_d = tl * K, tr_d = tr * K, br_d = br * K, bl_d = bl * K;
        const V2 tlc = v2(x_min, y_min), trc = v2(x_max, y_min), brc = v2(x_max, y_max), blc = v2(x_min, y_max);
        const V2 q[16] = {v2(x_min, y_min + tl), tlc + v2(0.0f, tl - tl_d), tlc + v2(tl - tl_d, 0.0f), tlc + v2(tl, 0.0f),
                          v2(x_max - tr, y_min), trc + v2(-tr + tr_d, 0.0f), trc + v2(0.0f, tr - tr_d), trc + v2(0.0f, tr),
                          v2(x_max, y_max - br), brc + v2(0.0f, -br + br_d), brc + v2(-br + br_d, 0.0f), brc + v2(-br, 0.0f),
                          v2(x_min + bl, y_max), blc + v2(bl - bl_d, 0.0f), blc + v2(0.0f, -bl + bl_d), blc + v2(0.0f, -bl)};
        if (positive) {
            o.begin(q[0]);
            if (tl > 0.0f) o.cubic_to(q[1], q[2], q[3]);
            o.line_to(q[4]);
            if (tr > 0.0f) o.cubic_to(q[5], q[6], q[7]);
            o.line_to(q[8]);
            if (br > 0.0f) o.cubic_to(q[9], q[10], q[11]);
            o.line_to(q[12]);
            if (bl > 0.0f) o.cubic_to(q[13], q[14], q[15]);
        } else {
            o.begin(q[15]);
            if (bl > 0.0f) o.cubic_to(q[14], q[13], q[12]);
            o.line_to(q[11]);
            if (br > 0.0f) o.cubic_to(q[10], q[9], q[8]);
            o.line_to(q[7]);
            if (tr > 0.0f) o.cubic_to(q[6], q[5], q[4]);
            o.line_to(q[3]);
            if (tl > 0.0f) o.cubic_to(q[2], q[1], q[0]);
        }
        o.end(true);
    }
    *n_verbs = o.nv;
    *n_points = o.np;
}

// StrokeBuilder::add_rectangle (stroke.rs:523-537) + approximate_thin_rectangle (stroke.rs:2336-2369): returns 1 and the
// stand-in line segment out[0..3], line width out[4], cap out[5] when the rectangle is approximated; 0 otherwise.
int oracle_thin_rectangle(const float* rect, float line_width, int line_join, float* out) {
    const float w = rect[2] - rect[0], h = rect[3] - rect[1];
    const float threshold = (line_join == 0 ? 1.0f : 0.05f) * line_width;
    if (!(fabsf(w) < threshold || fabsf(h) < threshold)) return 0;
    float d;
    if (w > h) {
        d = h * 0.5f;
        const float y = (rect[1] + rect[3]) * 0.5f;
        out[0] = rect[0] + d; out[1] = y; out[2] = rect[2] - d; out[3] = y;
    } else {
        d = w * 0.5f;
        const float x = (rect[0] + rect[2]) * 0.5f;
        out[0] = x; out[1] = rect[1] + d; out[2] = x; out[3] = rect[3] - d;
    }
    out[4] = line_width + d;
    out[5] = line_join == 2 ? 2.0f : 1.0f;  // LineJoin::Round -> LineCap::Round, else Square
    return 1;
}

// ---- small pure functions exposed for known-answer tests ----------------------------------------
void oracle_compute_normal(float v1x, float v1y, float v2x, float v2y, float* out) {
    V2 n = compute_normal(v2(v1x, v1y), v2(v2x, v2y));
    out[0] = n.x;
    out[1] = n.y;
}
float oracle_fast_atan2(float y, float x) { return fast_atan2(y, x); }
float oracle_angle_to(float a, float b) { return angle_to(a, b); }
float oracle_cubic_wang(const float* c, float tol) {
    return cubic_wang(v2(c[0], c[1]), v2(c[2], c[3]), v2(c[4], c[5]), v2(c[6], c[7]), tol);
}
float oracle_quad_wang(const float* c, float tol) { return quad_wang(v2(c[0], c[1]), v2(c[2], c[3]), v2(c[4], c[5]), tol); }
int oracle_find_sharp_turn(const float* c, float* t) {
    return find_sharp_turn(v2(c[0], c[1]), v2(c[2], c[3]), v2(c[4], c[5]), t) ? 1 : 0;
}
float oracle_circle_flattening_step(float r, float tol) { return circle_flattening_step(r, tol); }
int oracle_line_intersection(const float* a, const float* b, float* out) {
    V2 o = v2(0.0f, 0.0f);
    bool ok = line_intersection(line_d(v2(a[0], a[1]), v2(a[2], a[3])), line_d(v2(b[0], b[1]), v2(b[2], b[3])), &o);
    out[0] = o.x;
    out[1] = o.y;
    return ok ? 1 : 0;
}
// 0: transcendental calls evaluated in double and rounded once (default, what the device does); 1: the platform's f32 libm
void oracle_set_libm_mode(int platform_f32) { g_libm_platform_f32 = platform_f32 != 0; }
// libm trace (tests/test_libm_pinning.py): record every libm call of the calling thread between begin and end.
static LibmTrace g_trace_store;
void oracle_libm_trace_begin() {
    g_trace_store.fn.clear();
    g_trace_store.arg.clear();
    g_libm_trace = &g_trace_store;
}
uint64_t oracle_libm_trace_end() {
    g_libm_trace = nullptr;
    return g_trace_store.fn.size();
}
void oracle_libm_trace_copy(uint8_t* fn, float* arg) {
    if (!g_trace_store.fn.empty()) {
        memcpy(fn, g_trace_store.fn.data(), g_trace_store.fn.size());
        memcpy(arg, g_trace_store.arg.data(), g_trace_store.arg.size() * 4);
    }
}
}  // extern "C"
