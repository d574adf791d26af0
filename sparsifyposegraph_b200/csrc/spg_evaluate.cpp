// spg_evaluate.cpp — the replay loop the reference drives the removal path with:
//   evaluate()            reference src/evaluate.cpp:32-221   (incremental + baseline graphs grown vertex by vertex,
//                         decimation schedule, computeSubstituteEdge for links to already-marginalised vertices,
//                         optimize, marginalize, KLD / delta-chi2 samples, result-file layout)
//   parseLine()           reference src/main.cpp:9-103        (one job line of scripts/inputgenerator.sh)
// Host C++ over the graph container; every heavy step is a GPU call of this library: spg_graph_marginalize (the node
// removal path), spg_graph_optimize and spg_graph_kld / spg_graph_chi2 (spg_eval.cu).
#include <sys/stat.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <fstream>
#include <limits>
#include <set>
#include <sstream>
#include <string>
#include <vector>

#include "spg_host.h"

void spg_set_err(const std::string &s);

namespace {

using spg::Graph;
using spg::GraphEdge;

double nowS() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// GraphWrapperG2O::clonePortion (src/graph_wrapper_g2o.cpp:334-356): vertices and pose edges with ids <= maxid
spg_graph *clonePortion(const Graph &gw, int maxid) {
    spg_graph *out = new spg_graph;
    out->g = new Graph(gw.dim);
    for(int id : gw.vertexIds())
        if(id <= maxid) out->g->addVertex(id, gw.vertex(id)->pose);
    for(int ei : gw.edgeOrder()) {
        const GraphEdge &e = gw.edges[ei];
        if(e.kind != SPG_EDGE_POSE) continue;
        if(e.v[0] <= maxid && e.v[1] <= maxid) out->g->addPoseEdge(e.v[0], e.v[1], e.meas(), e.info());
    }
    return out;
}

// printStats (src/graph_wrapper_g2o.cpp:606-612): nodes (without the fixed one), edges, fill-in of the information matrix
void graphStats(const Graph &g, int fixed_id, int *nodes, int *edges, double *fillin) {
    std::vector<int> ids = g.vertexIds();
    *nodes = (int) ids.size() - 1;
    *edges = g.aliveEdges;
    std::set<std::pair<int, int>> pairs;
    for(int ei : g.edgeOrder()) {
        const GraphEdge &e = g.edges[ei];
        for(int a = 0; a < e.nv(); a++)
            for(int b = a + 1; b < e.nv(); b++) {
                if(e.v[a] == fixed_id || e.v[b] == fixed_id || e.v[a] == e.v[b]) continue;
                pairs.insert({std::min(e.v[a], e.v[b]), std::max(e.v[a], e.v[b])});
            }
    }
    const double n = std::max(1, *nodes);
    *fillin = (n + 2.0 * pairs.size()) / (n * n);
}

const char *kTopoShort[] = {"tree", "subgr", "clsubgr", "dense", "cldense"};
const char *kTopoLong[] = {"Tree", "Subgraph", "Cliquey Subgraph", "Dense", "Cliquey Dense"};
const char *kProfile[] = {"online", "cluster", "global"};

std::vector<int> decimate(const spg_evaluate_info &info, int last, int endvert) {
    spg::DecimateOptions o{info.sparsity, info.cluster_size};
    if(info.profile == SPG_PROFILE_ONLINE) return spg::onlineDecimate(last, endvert, o);
    if(info.profile == SPG_PROFILE_CLUSTER) return spg::clusterDecimate(last, endvert, o);
    return spg::globalDecimate(last, endvert, o);
}

} // namespace

extern "C" {

// parseLine (src/main.cpp:9-103): "<nfr|glc|none> <file.g2o> <online|cluster|global> <tree|subgr|clsubgr|dense|cldense>
// <local|global> <sparsity> [kldPeriod [chi2|kld [clusterSize]]]"
spg_status spg_evaluate_parse_job(const char *line, spg_evaluate_info *info, char *g2oname, int32_t cap) {
    if(!line || !info) return SPG_ERR_INVALID;
    std::stringstream ss(line);
    std::string tok;
    auto lower = [](std::string s) {
        std::transform(s.begin(), s.end(), s.begin(), ::tolower);
        return s;
    };
    *info = spg_evaluate_info{};
    if(!(ss >> tok)) return SPG_ERR_INVALID;
    tok = lower(tok);
    info->algorithm = tok == "glc" ? SPG_ALG_GLC : (tok == "none" ? SPG_EVAL_NONE : SPG_ALG_NFR);
    if(!(ss >> tok)) return SPG_ERR_INVALID;
    if(g2oname && cap > 0) {
        std::strncpy(g2oname, tok.c_str(), (size_t) cap - 1);
        g2oname[cap - 1] = 0;
        info->g2oname = g2oname;
    }
    ss >> tok;
    tok = lower(tok);
    info->profile = tok == "online" ? SPG_PROFILE_ONLINE : (tok == "cluster" ? SPG_PROFILE_CLUSTER : SPG_PROFILE_GLOBAL);
    ss >> tok;
    tok = lower(tok);
    info->opts.topology = tok == "tree" ? SPG_TOPO_TREE : tok == "subgr" ? SPG_TOPO_SUBGRAPH : tok == "clsubgr" ? SPG_TOPO_CLIQUEY_SUBGRAPH
                          : tok == "dense" ? SPG_TOPO_DENSE : SPG_TOPO_CLIQUEY_DENSE;
    ss >> tok;
    info->opts.lin_point = lower(tok) == "local" ? SPG_LIN_LOCAL : SPG_LIN_GLOBAL;
    info->opts.chord_ratio = 1.0;       // SparsityOptions defaults (src/sparsity_options.h:25-29)
    info->opts.include_intra_clique = 1;
    if(!(ss >> info->sparsity)) return SPG_ERR_INVALID;
    info->kld_period = 10;
    if(ss.good()) ss >> info->kld_period;
    if(info->profile == SPG_PROFILE_GLOBAL) info->kld_period = std::numeric_limits<int>::max();
    info->use_chi2 = 0;
    if(ss.good() && (ss >> tok)) info->use_chi2 = lower(tok) == "chi2";
    info->cluster_size = 100;
    if(ss.good()) ss >> info->cluster_size;
    return SPG_OK;
}

spg_status spg_evaluate(spg_ctx *ctx, const spg_graph *gwh, const spg_evaluate_info *info, int32_t *sample_vertex, double *sample_value,
                        int32_t cap, spg_evaluate_result *res, spg_graph **incremental_out, spg_graph **baseline_out) {
    if(!ctx || !gwh || !info || !res || info->sparsity <= 0 || info->kld_period <= 0 ||
       (info->profile == SPG_PROFILE_CLUSTER && info->cluster_size <= 0)) {
        spg_set_err("spg_evaluate: bad arguments");
        return SPG_ERR_INVALID;
    }
    const Graph &gw = *gwh->g;
    *res = spg_evaluate_result{};
    if(incremental_out) *incremental_out = nullptr;
    if(baseline_out) *baseline_out = nullptr;
    const int lastid = gw.maxVertexId();
    for(int i = 0; i <= std::min(lastid, 3); i++)
        if(!gw.hasVertex(i)) {
            spg_set_err("spg_evaluate: the graph needs vertices 0..3 (clonePortion(3))");
            return SPG_ERR_INVALID;
        }
    const int P = gw.poseWords(), dim = gw.dim;
    const int32_t fixed0 = 0;
    spg_graph *incremental = clonePortion(gw, 3), *baseline = clonePortion(gw, 3);
    auto fail = [&](spg_status s) {
        spg_graph_destroy(incremental);
        spg_graph_destroy(baseline);
        return s;
    };
    spg_status st = SPG_OK;
    auto optimize = [&](spg_graph *g) -> spg_status { // GraphWrapperG2O::optimize (:250-269)
        const double t = nowS();
        spg_status s = spg_graph_optimize(ctx, g, &fixed0, 1, 50, nullptr);
        res->seconds_optimize += nowS() - t;
        return s;
    };
    if((st = optimize(incremental)) != SPG_OK || (st = optimize(baseline)) != SPG_OK) return fail(st);

    std::set<int> marginalized;
    std::ofstream kldf, txtf;
    std::string savename;
    if(info->destdir) { // result-file layout of evaluate.cpp:69-96
        std::string g2o = info->g2oname ? info->g2oname : "graph";
        size_t a = g2o.rfind('/'), b = g2o.rfind('.');
        const size_t n0 = a == std::string::npos ? 0 : a + 1, n1 = (b == std::string::npos || b < n0) ? g2o.size() : b;
        const std::string dsname = g2o.substr(n0, n1 - n0);
        const std::string alg = info->algorithm == SPG_ALG_NFR ? (dim == 3 ? "se2" : "se3") : (info->algorithm == SPG_ALG_GLC ? "glc" : "none");
        std::string dir = info->destdir;
        mkdir(dir.c_str(), 0755);
        dir += std::string("/") + kProfile[info->profile];
        mkdir(dir.c_str(), 0755);
        dir += "/" + std::to_string(info->sparsity);
        mkdir(dir.c_str(), 0755);
        dir += "/" + dsname;
        mkdir(dir.c_str(), 0755);
        savename = dir + "/" + alg + "_" + kTopoShort[info->opts.topology] + "_" + (info->opts.lin_point == SPG_LIN_LOCAL ? "l" : "g");
        kldf.open(savename + ".kld");
        txtf.open(savename + ".txt");
        if(!kldf || !txtf) {
            spg_set_err("spg_evaluate: cannot write " + savename + ".kld");
            return fail(SPG_ERR_IO);
        }
        kldf.precision(6);
    }

    double kld = 0;
    std::vector<double> meas(P), infom((size_t) dim * dim);
    for(int i = 4; i <= lastid; i++) {
        const spg::GraphVertex *latest = gw.vertex(i);
        if(!latest) {
            spg_set_err("spg_evaluate: vertex " + std::to_string(i) + " is missing (ids must be contiguous)");
            return fail(SPG_ERR_INVALID);
        }
        spg_graph_add_vertex(incremental, i, latest->pose);
        spg_graph_add_vertex(baseline, i, latest->pose);
        // the edges of the new vertex towards older ones, in file order (evaluate.cpp:104-123)
        std::vector<int> es(latest->edges.begin(), latest->edges.end());
        std::sort(es.begin(), es.end(), [&](int x, int y) { return gw.edges[x].uidKey() < gw.edges[y].uidKey(); });
        for(int ei : es) {
            const GraphEdge &e = gw.edges[ei];
            if(e.kind != SPG_EDGE_POSE) continue;
            int from = e.v[0], to = e.v[1];
            if(from > i || to > i) continue;
            const int linkto = (from == i) ? to : from;
            if(marginalized.count(linkto) > 0) {
                spg::computeSubstituteEdge(&gw, marginalized, i, from, to, meas.data(), infom.data());
            } else {
                std::memcpy(meas.data(), e.meas(), sizeof(double) * P);
                std::memcpy(infom.data(), e.info(), sizeof(double) * dim * dim);
            }
            spg_graph_add_edge(incremental, from, to, meas.data(), infom.data());
            spg_graph_add_edge(baseline, from, to, meas.data(), infom.data());
        }
        std::vector<int> which = decimate(*info, i, lastid);
        const bool sample = (i % info->kld_period == 0) || i == lastid;
        if(info->algorithm != SPG_EVAL_NONE && (!which.empty() || sample)) {
            if((st = optimize(incremental)) != SPG_OK || (st = optimize(baseline)) != SPG_OK) return fail(st);
        }
        if(!which.empty() && info->algorithm != SPG_EVAL_NONE) {
            const double t = nowS();
            // GraphWrapper::marginalize = marginalizeNoOptimize + optimize (:455-463)
            st = spg_graph_marginalize(incremental, ctx, which.data(), (int32_t) which.size(), &info->opts, info->algorithm);
            res->seconds_marginalize += nowS() - t;
            if(st != SPG_OK) return fail(st);
            if((st = optimize(incremental)) != SPG_OK) return fail(st);
            marginalized.insert(which.begin(), which.end());
            res->n_marginalize_calls++;
            res->n_marginalized += (int32_t) which.size();
        } else if(info->algorithm == SPG_EVAL_NONE) {
            marginalized.insert(which.begin(), which.end());
        }
        if(sample) {
            const double t = nowS();
            if(info->algorithm == SPG_EVAL_NONE) {
                if((st = optimize(baseline)) != SPG_OK) return fail(st);
                kld = 0;
                if(info->use_chi2 && (st = spg_graph_chi2(ctx, baseline, &kld)) != SPG_OK) return fail(st);
            } else if(info->use_chi2) {
                // baseline->chi2(incremental) - baseline->chi2()  (:501-528): the baseline re-optimised with the
                // sparsified graph's vertices pinned at its estimates
                double c0 = 0, c1 = 0;
                if((st = spg_graph_chi2(ctx, baseline, &c0)) != SPG_OK) return fail(st);
                std::vector<int> bids = baseline->g->vertexIds();
                std::vector<double> saved((size_t) bids.size() * P);
                for(size_t q = 0; q < bids.size(); q++) std::memcpy(&saved[q * P], baseline->g->vertex(bids[q])->pose, sizeof(double) * P);
                std::vector<int32_t> pin;
                for(int id : incremental->g->vertexIds()) {
                    spg_graph_set_vertex_pose(baseline, id, incremental->g->vertex(id)->pose);
                    pin.push_back(id);
                }
                st = spg_graph_optimize(ctx, baseline, pin.data(), (int32_t) pin.size(), 50, nullptr);
                if(st == SPG_OK) st = spg_graph_chi2(ctx, baseline, &c1);
                for(size_t q = 0; q < bids.size(); q++) spg_graph_set_vertex_pose(baseline, bids[q], &saved[q * P]); // pop
                if(st != SPG_OK) return fail(st);
                kld = c1 - c0;
            } else {
                if((st = spg_graph_kld(ctx, baseline, incremental, fixed0, &kld, nullptr)) != SPG_OK) return fail(st);
            }
            res->seconds_kld += nowS() - t;
            if(kldf.is_open()) kldf << i << " " << kld << std::endl;
            if(res->n_samples < cap && sample_vertex && sample_value) {
                sample_vertex[res->n_samples] = i;
                sample_value[res->n_samples] = kld;
            }
            res->n_samples++;
        }
    }
    res->last_vertex = lastid;
    res->last_value = kld;
    graphStats(*baseline->g, fixed0, &res->baseline_nodes, &res->baseline_edges, &res->baseline_fillin);
    graphStats(*incremental->g, fixed0, &res->marginal_nodes, &res->marginal_edges, &res->marginal_fillin);
    if(txtf.is_open()) { // evaluate.cpp:196-203
        std::string alg = info->algorithm == SPG_ALG_NFR ? (dim == 3 ? "SE2" : "SE3") : (info->algorithm == SPG_ALG_GLC ? "GLC" : "NONE");
        txtf << alg << " " << kTopoLong[info->opts.topology];
        txtf << std::endl << "    baseline:     " << "nodes = " << res->baseline_nodes << "; edges = " << res->baseline_edges
             << "; fillin = " << res->baseline_fillin * 100 << "%";
        txtf << std::endl << "    marginalized: " << "nodes = " << res->marginal_nodes << "; edges = " << res->marginal_edges
             << "; fillin = " << res->marginal_fillin * 100 << "%";
        txtf << std::endl << "    last " << (info->use_chi2 ? "chi2: " : "kld: ") << kld << std::endl;
    }
    if(incremental_out) *incremental_out = incremental;
    else spg_graph_destroy(incremental);
    if(baseline_out) *baseline_out = baseline;
    else spg_graph_destroy(baseline);
    return SPG_OK;
}

} // extern "C"
