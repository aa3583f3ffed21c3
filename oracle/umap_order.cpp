// oracle/umap_order.cpp -- TEST INFRASTRUCTURE ONLY.
//
// The reference worker asks the parameter server for a sample's per-field tensors in the ITERATION ORDER of a
// std::unordered_map<size_t, size_t> that it clears and refills per sample (distributed_algo_abst.h:186-217,337:
// tensor_map; distribut/pull.h:72-80 walks it), and the server draws a tensor's initial values from its rand() stream the
// first time it sees the key (distribut/paramserver.h:40-46,146-151).  Which tensor gets which random numbers is therefore
// a property of libstdc++'s hash table.  This program replays exactly that sequence of clear() / insert() calls on the same
// container type, compiled by the same g++ as the reference roles, and prints the keys in the order the server first sees
// them; tests/golden/make_wnd_ref_curve.py stores the list next to the reference's loss curve.
//   usage: umap_order data.csv   (rows "label field:fid:val ...", the worker's own format, distributed_algo_abst.h:285-315)
#include <cstdio>
#include <fstream>
#include <set>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    std::ifstream fin(argv[1]);
    std::string line;
    std::unordered_map<size_t, size_t> tensor_map;  // the worker's member: lives across samples, cleared per sample
    std::unordered_set<size_t> seen;
    while (std::getline(fin, line)) {
        const char* p = line.c_str();
        int y, n;
        if (sscanf(p, "%d%n", &y, &n) < 1) continue;
        p += n + 1;
        size_t field, fid;
        float val;
        std::vector<std::pair<size_t, size_t>> row;
        while (p < line.c_str() + (int)line.length() && sscanf(p, "%zu:%zu:%f%n", &field, &fid, &val, &n) >= 2) {
            p += n + 1;
            row.emplace_back(fid, field);
        }
        if (row.empty()) continue;
        tensor_map.clear();
        std::set<size_t> fields;
        for (auto& e : row)
            if (fields.count(e.second) == 0) {
                tensor_map.insert(std::make_pair(e.first, e.second));
                fields.insert(e.second);
            }
        for (auto it = tensor_map.begin(); it != tensor_map.end(); ++it)
            if (seen.insert(it->first).second) printf("%zu\n", it->first);
    }
    return 0;
}
