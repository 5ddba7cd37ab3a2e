// Test driver for include/p2g.hpp.
//   host_mirror dump  < spec            -> JSON of everything CommonCircuitData derives (CPU only)
//   host_mirror prove spec cs.bin wires.bin out_prefix   -> proves through p2g::CircuitData (needs a GPU), writes
//                                                           <out_prefix>.proof and <out_prefix>.cproof (compressed)
// spec (text): "degree_bits num_public_inputs hasher num_wires" then one "kind p0 p1 p2 p3" line per gate, then for `prove` a
// line "pis v0 v1 ..." .
#include <cstdio>
#include <fstream>
#include <iostream>
#include <sstream>

#include "p2g.hpp"

static std::vector<uint64_t> read_bin(const char* path, size_t words) {
    std::vector<uint64_t> v(words);
    std::ifstream f(path, std::ios::binary);
    f.read((char*)v.data(), (std::streamsize)(words * 8));
    if ((size_t)f.gcount() != words * 8) throw std::runtime_error(std::string("short read: ") + path);
    return v;
}
static void write_bin(const std::string& path, const std::vector<uint8_t>& b) {
    std::ofstream f(path, std::ios::binary);
    f.write((const char*)b.data(), (std::streamsize)b.size());
}

int main(int argc, char** argv) {
    try {
        if (argc < 2) return 2;
        std::string mode = argv[1];
        std::ifstream file;
        if (mode == "prove") file.open(argv[2]);
        std::istream& in = mode == "prove" ? (std::istream&)file : std::cin;
        uint32_t degree_bits, npi, hasher, num_wires;
        in >> degree_bits >> npi >> hasher >> num_wires;
        p2g::CircuitConfig cfg;
        cfg.num_wires = num_wires;
        cfg.hasher = hasher;
        std::vector<p2g::Gate> gates;
        std::vector<uint64_t> pis;
        std::string line;
        std::getline(in, line);
        while (std::getline(in, line)) {
            if (line.empty()) continue;
            std::istringstream ls(line);
            if (line.rfind("pis", 0) == 0) {
                std::string tag;
                ls >> tag;
                uint64_t v;
                while (ls >> v) pis.push_back(v);
                continue;
            }
            uint32_t k, a, b, c, d;
            ls >> k >> a >> b >> c >> d;
            gates.emplace_back(k, a, b, c, d);
        }
        p2g::CommonCircuitData com(cfg, degree_bits, gates, npi);
        if (mode == "dump") {
            std::printf("{\"gates\": [");
            for (size_t i = 0; i < com.gates.size(); i++)
                std::printf("%s[%u, %u, %u, %u, %u]", i ? ", " : "", com.gates[i].kind, com.gates[i].params[0], com.gates[i].params[1],
                            com.gates[i].params[2], com.gates[i].params[3]);
            std::printf("], \"selector_indices\": [");
            for (size_t i = 0; i < com.selector_indices.size(); i++) std::printf("%s%u", i ? ", " : "", com.selector_indices[i]);
            std::printf("], \"groups\": [");
            for (size_t i = 0; i < com.groups.size(); i++) std::printf("%s[%u, %u]", i ? ", " : "", com.groups[i].first, com.groups[i].second);
            std::printf("], \"num_constants\": %u, \"num_gate_constraints\": %u, \"num_partial_products\": %u, \"num_selectors\": %u, ",
                        com.num_constants, com.num_gate_constraints, com.num_partial_products, com.num_selectors);
            std::printf("\"reduction_arity_bits\": [");
            for (size_t i = 0; i < com.reduction_arity_bits.size(); i++) std::printf("%s%u", i ? ", " : "", com.reduction_arity_bits[i]);
            std::printf("], \"k_is\": [");
            for (size_t i = 0; i < com.k_is.size(); i++) std::printf("%s%llu", i ? ", " : "", (unsigned long long)com.k_is[i]);
            std::printf("], \"ids\": [");
            for (size_t i = 0; i < com.gates.size(); i++) std::printf("%s\"%s\"", i ? ", " : "", com.gates[i].id().c_str());
            std::printf("]}\n");
            return 0;
        }
        if (mode == "prove" && argc >= 6) {
            const size_t n = com.degree();
            std::vector<uint64_t> cs = read_bin(argv[3], com.num_preprocessed() * n);
            std::vector<uint64_t> wires = read_bin(argv[4], (size_t)cfg.num_wires * n);
            p2g::CircuitData data(com, cs.data());
            p2g::ProofWithPublicInputs a = data.prove(wires.data(), pis);
            p2g::ProofWithPublicInputs b = data.prove(wires.data(), pis, nullptr, true);
            write_bin(std::string(argv[5]) + ".proof", a.to_bytes());
            write_bin(std::string(argv[5]) + ".cproof", b.to_bytes());
            // the witness as separately allocated columns (MatrixWitness.wire_values) and the VK file bytes
            std::vector<std::vector<uint64_t>> colv(cfg.num_wires);
            std::vector<const uint64_t*> colp(cfg.num_wires);
            for (uint32_t c = 0; c < cfg.num_wires; c++) {
                colv[c].assign(wires.begin() + (size_t)c * n, wires.begin() + (size_t)(c + 1) * n);
                colp[c] = colv[c].data();
            }
            if (data.prove_columns(colp, pis).to_bytes() != a.to_bytes()) return 5;
            write_bin(std::string(argv[5]) + ".vk", data.verifier_data_bytes());
            std::vector<uint8_t> cap;
            for (auto& c : data.constants_sigmas_cap) cap.insert(cap.end(), c.begin(), c.end());
            write_bin(std::string(argv[5]) + ".cap", cap);
            std::printf("ok %zu %zu launches %u total_ms %.3f\n", a.to_bytes().size(), b.to_bytes().size(), a.timings.kernel_launches,
                        a.timings.total_ms);
            // error behaviour: wrong public-input count throws (the reference unwrap()s)
            try {
                std::vector<uint64_t> bad(pis.size() + 1, 0);
                data.prove(wires.data(), bad);
                return 3;
            } catch (const p2g::Error& e) {
                if (e.code != P2G_EBADARG) return 4;
            }
            return 0;
        }
        return 2;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "host_mirror: %s\n", e.what());
        return 1;
    }
}
