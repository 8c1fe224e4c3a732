// warp_emul.cpp — TEST INFRASTRUCTURE ONLY: runs warp-uniform device code of star_b200/csrc/engine on the CPU.
//
// The seed-search code in seed_warp.cuh is written against the warp interface of warp_prims.cuh.  Here it is compiled with
// STAR_WARP_HOST_EMUL: the 32 lanes of a warp are 32 host threads that meet at a barrier in every collective (ballot, shuffle,
// reduction, syncwarp), so lane roles and the order of the collectives are exactly those of the GPU build.  tests/ compare the
// stored pieces of every read with the oracle's (star_oracle_map_chunk_dump).  Nothing in the product loads this file.
#define STAR_WARP_HOST_EMUL 1
#include <pthread.h>

#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "../star_b200/csrc/engine/seed_warp.cuh"

using namespace starb;

namespace {

struct HostIndex {   // DevIndex over host memory: aligned 64-bit copies of the packed arrays (+2 words), genome with its padding
    DevIndex ix;
    std::vector<u64> sa, sai;
    explicit HostIndex(const star_index_view_t* v) {
        memset(&ix, 0, sizeof(ix));
        ix.G = v->G;                       // the loader keeps >= 256 bytes of code 5 on both sides
        ix.nGenome = v->nGenome;
        sa.assign((v->nSAbyte + 7) / 8 + 2, 0);
        memcpy(sa.data(), v->SA, v->nSAbyte);
        sai.assign((v->nSAibyte + 7) / 8 + 2, 0);
        memcpy(sai.data(), v->SAi, v->nSAibyte);
        ix.SA = sa.data(); ix.nSA = v->nSA;
        ix.SAi = sai.data(); ix.nSAi = v->nSAi;
        ix.GstrandBit = v->GstrandBit;
        ix.saBits = v->GstrandBit + 1; ix.saiBits = v->GstrandBit + 3;
        ix.gSAindexNbases = v->gSAindexNbases;
        ix.GstrandMask = ~(1ULL << v->GstrandBit);
        ix.SAiMarkNmaskC = 1ULL << (v->GstrandBit + 1);
        ix.SAiMarkNmask = ~ix.SAiMarkNmaskC;
        ix.SAiMarkAbsentMaskC = 1ULL << (v->GstrandBit + 2);
        for (int i = 0; i < 20 && i <= (int)v->gSAindexNbases; i++) ix.genomeSAindexStart[i] = v->genomeSAindexStart[i];
    }
};

inline u8 nt2num(char c) {
    switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3; default: return 4; }
}

}  // namespace

extern "C" {
#pragma GCC visibility push(default)

// Seeds every read of the batch with the emulated warp.  pc receives 8 numbers per stored piece in the oracle's dump layout
// (rStart, Length, Str=0, Dir, Nrep, SAstart, SAend, iFrag); pcOff[nReads+1]; perRead[4*i..] = nP, nA, flags, multNminL.
// counters[4] += searches, SAi words, probes, rounds.  Returns 0, or 1 when pcCap is too small.
int warp_emul_seed_chunk(const star_index_view_t* view, const star_params_t* params, const star_read_batch_t* in, uint64_t* pcOff, uint64_t* pc,
                         uint64_t pcCap, uint32_t* perRead, uint64_t* counters) {
    HostIndex hix(view);
    const star_params_t P = *params;
    const uint32_t n = in->nReads;
    HostWarpShared shared;
    std::vector<u8> Rbuf(STAR_READ_SEQ_LENGTH_MAX + 64, 0);   // 16 bytes of slack on both sides (8-byte gathers of the compare)
    u8* R = Rbuf.data() + 16;
    std::vector<Piece> slab(P.seedPerReadNmax + 2);
    SeedWarpOut outs[32];
    uint32_t Lread = 0;
    int rcAll = 0;
    uint64_t nPieces = 0;
    pcOff[0] = 0;
    // 32 persistent lane threads; the read loop itself is uniform (every lane iterates the same reads)
    auto laneMain = [&](unsigned lane) {
        HostWarp w(lane, &shared);
        for (uint32_t i = 0; i < n; i++) {
            if (lane == 0) {   // prep_reads_kernel: Read1[0] = mate1 | spacer | revcomp(mate2)
                const uint64_t* off = in->seqOff + (uint64_t)i * in->nMates;
                const uint32_t l0 = (uint32_t)(off[1] - off[0]);
                const uint32_t l1 = in->nMates == 2 ? (uint32_t)(off[2] - off[1]) : 0;
                for (uint32_t k = 0; k < l0; k++) R[k] = nt2num(in->seq[off[0] + k]);
                Lread = l0;
                if (in->nMates == 2) {
                    R[l0] = STAR_MARK_FRAG_SPACER_BASE;
                    for (uint32_t k = 0; k < l1; k++) { u8 c = nt2num(in->seq[off[1] + l1 - 1 - k]); R[l0 + 1 + k] = c < 4 ? 3 - c : c; }
                    Lread = l0 + l1 + 1;
                }
            }
            w.sync();
            SeedWarpOut& st = outs[lane];
            st.PC = slab.data(); st.maxP = (u32)slab.size();
            warpSeedRead<HostWarp>(w, hix.ix, P, R, Lread, st);
            w.sync();
            if (lane == 0) {
                for (int l = 1; l < 32; l++)   // the lanes must agree on every uniform value
                    if (outs[l].nP != st.nP || outs[l].nA != st.nA || outs[l].flags != st.flags || outs[l].searches != st.searches) rcAll = 2;
                if (nPieces + st.nP > pcCap) rcAll = 1;
                else {
                    for (u32 k = 0; k < st.nP; k++) {
                        const Piece& p = st.PC[k];
                        uint64_t* o = pc + (nPieces + k) * 8;
                        o[0] = p.rStart; o[1] = p.Length; o[2] = 0; o[3] = p.Dir; o[4] = p.Nrep; o[5] = p.SAstart; o[6] = p.SAstart + p.Nrep - 1; o[7] = p.iFrag;
                    }
                    nPieces += st.nP;
                }
                pcOff[i + 1] = nPieces;
                perRead[4 * i] = st.nP; perRead[4 * i + 1] = st.nA; perRead[4 * i + 2] = st.flags; perRead[4 * i + 3] = st.multNminL;
                counters[0] += st.searches; counters[1] += st.saiWords; counters[2] += st.probes; counters[3] += st.rounds;
            }
            w.sync();
        }
    };
    std::vector<std::thread> th;
    for (unsigned l = 0; l < 32; l++) th.emplace_back(laneMain, l);
    for (auto& t : th) t.join();
    return rcAll;
}

#pragma GCC visibility pop
}  // extern "C"
