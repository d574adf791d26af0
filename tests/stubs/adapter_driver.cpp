// CPU end-to-end check of integration/spg_vertex_remover_g2o.hpp against the stand-in g2o / reference interfaces of
// this directory (tests/test_g2o_adapter.py compiles and runs it). The device is replaced by an Engine that walks the
// library's own wavefront rounds (spg_graph_rounds_begin / round_next / round_apply) with the CPU oracle computing the
// blankets — so everything but the kernels is the product's code. The same removal is run directly on an spg_graph; the
// g2o graph the adapter leaves behind must hold exactly the same vertices and factors.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "spg_vertex_remover_g2o.hpp"

extern "C" double orc_remove_round(const spg_round_in *in, spg_round_out *out, int n_threads); // oracle/capi.cpp (checker)

static spg_status cpuEngine(spg_graph *g, const int32_t *which, int32_t n, const spg_sparsity_options *o, int32_t alg) {
    spg_status st = spg_graph_rounds_begin(g, which, n, o, alg);
    if(st != SPG_OK) return st;
    for(;;) {
        spg_round_in r;
        st = spg_graph_round_next(g, &r);
        if(st != SPG_OK) return st;
        if(r.n_blankets == 0) return SPG_OK;
        std::vector<uint64_t> out((size_t) r.out_off[r.n_blankets]);
        spg_round_out ro{out.data(), NULL, NULL, NULL, NULL};
        orc_remove_round(&r, &ro, 1);
        st = spg_graph_round_apply(g, out.data());
        if(st != SPG_OK) return st;
    }
}

static std::mt19937 rng(12345);
static double uni(double a, double b) { return a + (b - a) * (rng() / 4294967296.0); }

static void randomInfo(int d, double *info) { // SPD, column-major
    std::vector<double> A((size_t) d * d);
    for(double &x : A) x = uni(-1, 1);
    for(int i = 0; i < d; i++)
        for(int j = 0; j < d; j++) {
            double s = i == j ? 2.0 : 0.0;
            for(int k = 0; k < d; k++) s += A[i + k * d] * A[j + k * d];
            info[i + j * d] = 10.0 * s;
        }
}

struct Factor { int from, to; double meas[7]; double info[36]; };

static int fail(const char *what) { std::printf("FAIL: %s\n", what); return 1; }

static int runCase(int dim, bool glc, SparsityOptions::SparsityTopology topo) {
    const int P = dim == 3 ? 3 : 7, N = 40;
    // a chain with loop closures, noiseless measurements z = x_i^-1 x_j of a planar / yaw-only trajectory
    std::vector<std::vector<double>> pose(N, std::vector<double>(7, 0.0));
    std::vector<double> yaw(N);
    for(int i = 0; i < N; i++) {
        yaw[i] = 0.15 * i;
        pose[i][0] = 2.0 * std::cos(0.2 * i) + 0.1 * i;
        pose[i][1] = 2.0 * std::sin(0.2 * i);
        if(dim == 3) pose[i][2] = yaw[i];
        else { pose[i][2] = 0.05 * i; pose[i][3] = 0; pose[i][4] = 0; pose[i][5] = std::sin(yaw[i] / 2); pose[i][6] = std::cos(yaw[i] / 2); }
    }
    auto relative = [&](int i, int j, double *z) {
        const double c = std::cos(yaw[i]), s = std::sin(yaw[i]);
        const double dx = pose[j][0] - pose[i][0], dy = pose[j][1] - pose[i][1];
        z[0] = c * dx + s * dy + uni(-0.01, 0.01);
        z[1] = -s * dx + c * dy + uni(-0.01, 0.01);
        const double dth = yaw[j] - yaw[i];
        if(dim == 3) z[2] = dth;
        else { z[2] = pose[j][2] - pose[i][2]; z[3] = 0; z[4] = 0; z[5] = std::sin(dth / 2); z[6] = std::cos(dth / 2); }
    };
    std::vector<Factor> factors;
    for(int i = 0; i + 1 < N; i++) { Factor f; f.from = i; f.to = i + 1; relative(i, i + 1, f.meas); randomInfo(dim, f.info); factors.push_back(f); }
    for(int i = 0; i + 7 < N; i += 3) { Factor f; f.from = i; f.to = i + 7; relative(i, i + 7, f.meas); randomInfo(dim, f.info); factors.push_back(f); }

    // ---- the g2o side -------------------------------------------------------------------------------------------------
    g2o::SparseOptimizer so;
    for(int i = 0; i < N; i++) {
        g2o::OptimizableGraph::Vertex *v = dim == 3 ? static_cast<g2o::OptimizableGraph::Vertex *>(new g2o::VertexSE2)
                                                    : static_cast<g2o::OptimizableGraph::Vertex *>(new g2o::VertexSE3);
        v->setId(i);
        v->setEstimateData(pose[i].data());
        so.addVertex(v);
    }
    SpgVertexRemover::EdgeMap edgeLookup;
    for(const Factor &f : factors) {
        g2o::OptimizableGraph::Edge *e = dim == 3 ? static_cast<g2o::OptimizableGraph::Edge *>(new EdgeSE2ISAM)
                                                  : static_cast<g2o::OptimizableGraph::Edge *>(new EdgeSE3ISAM);
        e->setVertex(0, so.vertex(f.from));
        e->setVertex(1, so.vertex(f.to));
        e->setMeasurementData(f.meas);
        for(int q = 0; q < dim * dim; q++) e->informationData()[q] = f.info[q];
        so.addEdge(e);
        edgeLookup[e] = new GraphWrapper::Edge;
    }
    // ---- the same graph directly in the library (factors in the adapter's mirror order: by vertex ids) ------------------
    spg_graph *ref = NULL;
    if(spg_graph_create(&ref, dim) != SPG_OK) return fail("spg_graph_create");
    for(int i = 0; i < N; i++) spg_graph_add_vertex(ref, i, pose[i].data());
    std::vector<Factor> sorted = factors;
    std::stable_sort(sorted.begin(), sorted.end(), [](const Factor &a, const Factor &b) { return a.from != b.from ? a.from < b.from : a.to < b.to; });
    for(const Factor &f : sorted) spg_graph_add_edge(ref, f.from, f.to, f.meas, f.info);

    SparsityOptions opts;
    opts.topology = topo;
    opts.linPoint = SparsityOptions::Global;
    spg_sparsity_options o{(int32_t) topo, SPG_LIN_GLOBAL, 1.0, 1, 0};
    // two calls, like successive evaluate() steps: the second one meets the factors the first one created
    const std::vector<std::vector<int>> lists = {{3, 8, 13, 20, 21, 30}, {5, 14, 22, 31, 9}};
    for(const std::vector<int> &list : lists) {
        SpgVertexRemover vr(NULL);
        if(glc) vr.registerTopologyProvider(new TopologyProviderGLC);
        else { vr.registerTopologyProvider(new TopologyProviderSE2ISAM); vr.registerTopologyProvider(new TopologyProviderSE3ISAM); }
        vr.setGraph(&so);
        vr.setEdgeMap(&edgeLookup);
        vr.setSparsityOptions(opts);
        vr.setEngine(cpuEngine);
        std::vector<g2o::OptimizableGraph::Vertex *> removeList;
        std::vector<int32_t> which;
        for(int id : list) { removeList.push_back(so.vertex(id)); which.push_back(id); }
        const size_t edgesBefore = so.edges().size();
        const int erasedBefore = edgeLookup.erased;
        g2o::OptimizableGraph::EdgeContainer added = vr.remove(removeList);
        for(g2o::OptimizableGraph::Edge *e : added) edgeLookup[e] = new GraphWrapper::Edge; // src/graph_wrapper_g2o.cpp:448-450
        if(cpuEngine(ref, which.data(), (int32_t) which.size(), &o, glc ? SPG_ALG_GLC : SPG_ALG_NFR) != SPG_OK) return fail("reference removal");
        // ---- compare ---------------------------------------------------------------------------------------------------
        if((int) so.vertices().size() != spg_graph_num_vertices(ref)) return fail("vertex count");
        for(int id : list) if(so.vertex(id)) return fail("removed vertex still in g2o");
        if((int) so.edges().size() != spg_graph_num_edges(ref)) return fail("edge count");
        if(edgeLookup.size() != so.edges().size()) return fail("edge map out of step with the graph");
        const int removedEdges = (int) edgesBefore + (int) added.size() - (int) so.edges().size();
        if(edgeLookup.erased - erasedBefore != removedEdges) return fail("edge map erase calls");
        // (Dense topologies remove several listed vertices with one extended blanket: count removals, not blankets)
        if(vr.lastStats().n_failed != 0 || vr.lastStats().n_applied != (int) list.size()) return fail("stats");
        // every factor of the library graph must be in g2o with the same payload
        const int ne = spg_graph_num_edges(ref);
        for(int i = 0; i < ne; i++) {
            spg_edge_desc d;
            spg_graph_edge_desc(ref, i, &d);
            std::vector<int32_t> ids(d.nv);
            const int cols = dim * d.nv, nm = d.kind == SPG_EDGE_MULTI ? d.rows / dim : 1;
            std::vector<double> meas(d.kind == SPG_EDGE_GLC ? cols : nm * P), info(d.kind == SPG_EDGE_GLC ? (size_t) d.rows * cols : (size_t) d.rows * d.rows);
            spg_graph_edge_data(ref, i, ids.data(), meas.data(), info.data());
            bool found = false;
            for(g2o::HyperGraph::Edge *he : so.edges()) {
                if((int) he->vertices().size() != d.nv) continue;
                bool same = true;
                for(int q = 0; q < d.nv; q++) same = same && he->vertices()[q] && he->vertices()[q]->id() == ids[q];
                if(!same) continue;
                const g2o::OptimizableGraph::Edge *e = static_cast<const g2o::OptimizableGraph::Edge *>(he);
                if(e->dimension() != d.rows) continue;
                // In the second call the mirror numbers every factor as a file edge (sorted by vertex ids) while the direct
                // run keeps the creation order of the first call's substitutes: the blankets sum their edges in a
                // different order, so payloads agree to rounding, not bit for bit. GLC weights through W^T W (the signs
                // of the eigenvectors are free).
                double worst = 0, scale = 1e-300;
                if(const GLCEdge *ge = dynamic_cast<const GLCEdge *>(he)) {
                    if(d.kind != SPG_EDGE_GLC) continue;
                    for(int a = 0; a < cols; a++)
                        for(int c = 0; c < cols; c++) {
                            double x = 0, y = 0;
                            for(int r = 0; r < d.rows; r++) {
                                x += ge->linearWeight()(r, a) * ge->linearWeight()(r, c);
                                y += info[(size_t) r * cols + a] * info[(size_t) r * cols + c];
                            }
                            worst = std::max(worst, std::fabs(x - y));
                            scale = std::max(scale, std::fabs(y));
                        }
                    for(int c = 0; c < cols; c++) worst = std::max(worst, scale * std::fabs(ge->measurement()(c) - meas[c]));
                    if(!dynamic_cast<const GLCReparamSE2ISAM *>(ge->reparam()) && !dynamic_cast<const GLCReparamSE3 *>(ge->reparam())) return fail("GLC reparam");
                    if(ge->information()(0, 0) != 1.0) return fail("GLC information");
                } else if(d.kind == SPG_EDGE_POSE) {
                    double z[7];
                    e->getMeasurementData(z);
                    for(int q = 0; q < dim * dim; q++) {
                        worst = std::max(worst, std::fabs(e->informationData()[q] - info[q]));
                        scale = std::max(scale, std::fabs(info[q]));
                    }
                    for(int q = 0; q < P; q++) worst = std::max(worst, scale * std::fabs(z[q] - meas[q]));
                } else {
                    for(size_t q = 0; q < info.size(); q++) {
                        worst = std::max(worst, std::fabs(e->informationData()[q] - info[q]));
                        scale = std::max(scale, std::fabs(info[q]));
                    }
                }
                if(worst <= 1e-8 * scale) { found = true; break; }
            }
            if(!found) { std::printf("factor %d (kind %d, %d vertices, first %d) has no twin in g2o\n", i, d.kind, d.nv, ids[0]); return fail("factor missing"); }
        }
    }
    spg_graph_destroy(ref);
    std::printf("ok: dim %d %s topology %d: %zu vertices, %zu factors left in g2o\n", dim, glc ? "GLC" : "NFR", (int) topo, so.vertices().size(), so.edges().size());
    return 0;
}

int main() {
    int bad = 0;
    bad += runCase(3, false, SparsityOptions::Tree);
    bad += runCase(6, false, SparsityOptions::Tree);
    bad += runCase(3, true, SparsityOptions::Tree);
    bad += runCase(6, true, SparsityOptions::Dense);
    bad += runCase(3, false, SparsityOptions::CliqueySubgraph);
    bad += runCase(6, false, SparsityOptions::CliqueyDense);
    return bad ? 1 : 0;
}
