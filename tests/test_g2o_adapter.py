"""The g2o adapter of INTEGRATION.md (integration/spg_vertex_remover_g2o.hpp: the reference's VertexRemover interface,
src/vertex_remover.h:19-50, on top of the C ABI) is compiled here against stand-ins of the g2o / reference interfaces it
touches (tests/stubs/ — g2o is not installed in this environment) and run end to end on the CPU: the product's scheduler
walks the rounds, the oracle computes the blankets, and the g2o graph the adapter leaves behind (removeEdge / edge-map
erase / removeVertex / addEdge, updateInputGraph :500-546) must hold the same vertices and factors as the same removal
run directly on an spg_graph — NFR tree (SE2, SE3), GLC tree / dense and the correlated topologies, two successive calls
each so that GLCEdge / MultiEdgeCorrelated factors also travel back into the library."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_adapter_compiles_against_the_interfaces_and_mirrors_the_graph(oracle, tmp_path):
    from sparsifyposegraph_b200 import capi
    capi.lib()                                   # the library is built
    stubs = os.path.join(ROOT, "tests", "stubs")
    libdir = os.path.join(ROOT, "sparsifyposegraph_b200")
    orcdir = os.path.join(ROOT, "oracle", "_build")
    exe = str(tmp_path / "adapter_driver")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-std=c++17", "-O1", "-Wall", "-Werror", "-I", stubs, "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(ROOT, "integration"), "-o", exe, os.path.join(stubs, "adapter_driver.cpp"),
                           "-L", libdir, "-lspg_b200", "-L", orcdir, "-lspg_oracle",
                           "-Wl,-rpath," + libdir, "-Wl,-rpath," + orcdir])
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(res.stdout)
    assert res.returncode == 0, res.stdout + res.stderr
    assert res.stdout.count("ok:") == 6
