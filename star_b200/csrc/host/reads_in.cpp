// reads_in.cpp — FASTQ/FASTA input, chunked.
//
// Restates the text handling of ReadAlignChunk::processChunks (reference
// source/ReadAlignChunk_processChunks.cpp:111-196: read ID = first token of line 1, Illumina filter flag from
// the 2nd token, sequence / '+' / quality lines with one trailing control character removed) and of readLoad
// (readLoad.cpp:4-100: name cut at --readNameSeparator, length checks).  Names and qualities stay on the host;
// only the sequences go to the engine.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <cstdlib>
#include <cstring>
#include <thread>

#include "host.h"

namespace starhost {

ReadsReader::~ReadsReader() { closeFiles(); }

// Maps a plain file read-only; returns false when the fast path does not apply (empty file, not a regular file, mmap failure).
static bool mapFile(const std::string& path, const char*& ptr, size_t& size) {
    int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) return false;
    struct stat st;
    if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode) || st.st_size == 0) { ::close(fd); return false; }
    void* p = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    ::close(fd);
    if (p == MAP_FAILED) return false;
    madvise(p, (size_t)st.st_size, MADV_SEQUENTIAL);
    ptr = (const char*)p; size = (size_t)st.st_size;
    return true;
}

// start of line number `line` (0-based) counted from `off`; = size when the file has fewer lines
static size_t skipLines(const char* base, size_t size, size_t off, uint64_t line) {
    while (line > 0 && off < size) {
        const char* q = (const char*)memchr(base + off, '\n', size - off);
        if (!q) return size;
        off = (size_t)(q - base) + 1;
        line--;
    }
    return off;
}

void ReadsReader::closeFiles() {
    for (int m = 0; m < 2; m++) {
        if (f[m]) { if (piped[m]) pclose(f[m]); else fclose(f[m]); f[m] = nullptr; }
        piped[m] = false;
        if (map[m] && mapSize[m]) munmap((void*)map[m], mapSize[m]);
        map[m] = nullptr; mapSize[m] = 0; mapOff[m] = 0;
        bpos[m] = blen[m] = 0;
    }
    fast = false;
}

// Opens file `idx` of the --readFilesIn lists (Parameters_openReadsFiles.cpp:5-105; the reference concatenates the files of a list
// through a FIFO with "FILE n" marker lines, here they are simply opened one after the other).
int ReadsReader::openFile(uint32_t idx, std::string& err) {
    closeFiles();
    fileIdx = idx;
    const HostParams& Pin = *P;
    if (Pin.readFilesCommand[0] == "-") {   // plain files: 4-line FASTQ goes through the memory-mapped parallel parser
        bool ok = true;
        for (unsigned m = 0; m < nMates && ok; m++) ok = mapFile(Pin.readFilesNames[m][idx], map[m], mapSize[m]);
        ok = ok && map[0][0] == '@';
        if (ok) { fast = true; return 0; }
        for (unsigned m = 0; m < nMates; m++) { if (map[m] && mapSize[m]) munmap((void*)map[m], mapSize[m]); map[m] = nullptr; mapSize[m] = 0; }
    }
    for (unsigned m = 0; m < nMates; m++) {
        const std::string& name = Pin.readFilesNames[m][idx];
        if (Pin.readFilesCommand[0] != "-") {  // Parameters_openReadsFiles.cpp:83-101 pipes the command's stdout
            std::string cmd;
            for (auto& w : Pin.readFilesCommand) cmd += w + " ";
            cmd += "\"" + name + "\"";
            f[m] = popen(cmd.c_str(), "r");
            piped[m] = true;
        } else {
            f[m] = fopen(name.c_str(), "rb");
        }
        if (!f[m]) {
            err = "EXITING because of fatal input ERROR: could not open readFilesIn=" + name + "\n";
            return STAR_EXIT_INPUT_FILES;
        }
        buf[m].resize(1 << 22);
        bpos[m] = blen[m] = 0;
    }
    return 0;
}

int ReadsReader::open(const HostParams& Pin, std::string& err) {
    P = &Pin;
    nMates = Pin.readNmates;
    const uint32_t nFiles = (uint32_t)Pin.readFilesNames[0].size();
    if (Pin.gpuShardCount > 1) {
        // contiguous slice of the input by read index: count the records of mate 1 in every file first (one pass over the text)
        std::vector<uint64_t> recs(nFiles, 0);
        uint64_t nRec = 0;
        for (uint32_t fi = 0; fi < nFiles; fi++) {
            int rc = openFile(fi, err);
            if (rc) return rc;
            if (fast) {
                uint64_t lines = 0, lastText = 0;   // lastText: number of lines up to the last non-empty one (trailing blank lines end the input)
                for (size_t off = 0; off < mapSize[0];) {
                    const char* q = (const char*)memchr(map[0] + off, '\n', mapSize[0] - off);
                    lines++;
                    const size_t len = q ? (size_t)(q - map[0]) - off : mapSize[0] - off;
                    if (len > 0 && !(len == 1 && map[0][off] == '\r')) lastText = lines;
                    if (!q) break;
                    off = (size_t)(q - map[0]) + 1;
                }
                recs[fi] = (lastText + 3) / 4;   // the same rule as nextFast: a last record with missing lines is still a record
            } else {
                std::string line;
                const bool fq = peekChar(0) == '@';
                uint64_t lines = 0, nr = 0;
                while (getLine(0, line)) { if (fq) lines++; else if (!line.empty() && line[0] == '>') nr++; }
                recs[fi] = fq ? lines / 4 : nr;
            }
            nRec += recs[fi];
        }
        shardLo = nRec * Pin.gpuShardIndex / Pin.gpuShardCount;
        shardHi = nRec * (Pin.gpuShardIndex + 1) / Pin.gpuShardCount;
        // position at record shardLo: skip whole files, then records inside the file (both mates), keeping the global read numbering
        uint64_t before = 0;
        uint32_t fi = 0;
        while (fi + 1 < nFiles && before + recs[fi] <= shardLo) { before += recs[fi]; fi++; }
        int rc = openFile(fi, err);
        if (rc) return rc;
        iReadAll = before;
        if (fast) {
            for (unsigned m = 0; m < nMates; m++) mapOff[m] = skipLines(map[m], mapSize[m], 0, 4 * (shardLo - before));
            iReadAll = shardLo;
        } else {
            std::string tmp;
            while (iReadAll < shardLo) {
                int ch = peekChar(0);
                if (ch != '@' && ch != '>') break;
                const bool fq = ch == '@';
                for (unsigned m = 0; m < nMates; m++) {
                    getLine(m, tmp);
                    if (fq) { getLine(m, tmp); getLine(m, tmp); getLine(m, tmp); }
                    else { for (;;) { int c2 = peekChar(m); if (c2 == '@' || c2 == '>' || c2 == ' ' || c2 == '\n' || c2 < 0) break; getLine(m, tmp); } }
                }
                iReadAll++;
            }
        }
        return 0;
    }
    return openFile(0, err);
}

int ReadsReader::peekChar(int m) {
    if (bpos[m] == blen[m]) {
        blen[m] = fread(buf[m].data(), 1, buf[m].size(), f[m]);
        bpos[m] = 0;
        if (blen[m] == 0) return -1;
    }
    return (unsigned char)buf[m][bpos[m]];
}

bool ReadsReader::getLine(int m, std::string& line) {
    line.clear();
    bool any = false;
    for (;;) {
        if (bpos[m] == blen[m]) {
            blen[m] = fread(buf[m].data(), 1, buf[m].size(), f[m]);
            bpos[m] = 0;
            if (blen[m] == 0) return any;
        }
        any = true;
        char* s = buf[m].data() + bpos[m];
        char* e = (char*)memchr(s, '\n', blen[m] - bpos[m]);
        if (e) {
            line.append(s, e - s);
            bpos[m] += (e - s) + 1;
            return true;
        }
        line.append(s, blen[m] - bpos[m]);
        bpos[m] = blen[m];
    }
}

static inline void stripEndControl(std::string& s) {  // fastqReadOneLine :284-297, removeStringEndControl :299-303
    if (!s.empty() && (int)(signed char)s.back() < 33) s.pop_back();
}

// One FASTQ record from its line pointers: ls/le[4*m + k] = start / end (newline excluded) of line k of mate m; a missing line has
// ls == nullptr.  Same text rules and error messages as the stream parser below.
// ClipMate::clip for the 5' end (fixed number of bases) and then the 3' end (fixed number, adapter by localSearch, bases after the
// adapter) of one mate (ClipMate_clip.cpp:5-78, SequenceFuns.cpp:293-315); appends the clip amounts and the clipped sequence for the engine
static void clipMateAppend(const HostParams& P, unsigned m, const char* s, uint64_t L0, ReadChunk& c) {
    uint64_t L = L0, c5 = 0, c3 = 0;
    if (P.clip5N[m] > 0) { if (L > P.clip5N[m]) { L -= P.clip5N[m]; c5 = P.clip5N[m]; } else { c5 = L; L = 0; } }
    const uint64_t Lold = L;
    if (P.clip3N[m] > 0) { if (L > P.clip3N[m]) { L -= P.clip3N[m]; c3 += P.clip3N[m]; } else { L = 0; c3 = Lold; } }
    if (!P.clip3Ad[m].empty()) {
        const std::string& ad = P.clip3Ad[m];
        auto num = [](char ch) -> int { switch (ch) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3; default: return 4; } };
        uint64_t nMatchBest = 0, nMMbest = 0, ixBest = L;
        for (uint64_t ix = 0; ix < L; ix++) {
            uint64_t nMatch = 0, nMM = 0;
            const uint64_t ny = std::min<uint64_t>(ad.size(), L - ix);
            for (uint64_t iy = 0; iy < ny; iy++) {
                const int x = num(s[c5 + ix + iy]);
                if (x > 3) continue;
                if (x == ad[iy]) nMatch++; else nMM++;
            }
            if ((nMatch > nMatchBest || (nMatch == nMatchBest && nMM < nMMbest)) && double(nMM) / double(nMatch) <= P.clip3MMp[m]) { ixBest = ix; nMatchBest = nMatch; nMMbest = nMM; }
        }
        const uint64_t clippedAdN = L - ixBest;
        L -= clippedAdN; c3 += clippedAdN;
    }
    if (P.clip3After[m] > 0 && (P.clip3N[m] > 0 || !P.clip3Ad[m].empty())) {   // (ClipMate_initialize.cpp:22-23: an end with neither N nor adapter is not clipped at all)
        if (L > P.clip3After[m]) { L -= P.clip3After[m]; c3 += P.clip3After[m]; } else { L = 0; c3 = Lold; } }
    c.clip5.push_back((uint16_t)c5);
    c.clip3.push_back((uint16_t)(L0 - c5 - L));
    if (L == 0 && c.nMates == 1) c.seqC.push_back('N'); else c.seqC.append(s + c5, L);   // (an empty single-end read is passed as one N: it maps nowhere either)
    c.seqOffC.push_back(c.seqC.size());
}

int ReadsReader::parseRecord(ReadChunk& c, uint64_t iRead, const char* const* ls, const char* const* le, std::string& err) const {
    auto strip = [](const char* s, const char*& e) { if (e > s && (int)(signed char)e[-1] < 33) e--; };   // removeStringEndControl
    const char* s0 = ls[0];
    const char* e0 = le[0];
    const char* t = s0;
    while (t < e0 && !isspace((unsigned char)*t)) t++;
    const char* idEnd = t;
    strip(s0, idEnd);
    char passFilter = 'N';
    {
        const char* q0 = t;
        while (q0 < e0 && isspace((unsigned char)*q0)) q0++;
        const char* q1 = q0;
        while (q1 < e0 && !isspace((unsigned char)*q1)) q1++;
        if (t < e0 && q1 - q0 >= 3 && q0[1] == ':' && q0[2] == 'Y' && (q1 - q0 > 3 ? q0[3] : '\0') == ':') passFilter = 'Y';
    }
    std::string readID = P->outSAMreadID == "Number" ? std::string("@") + std::to_string(iRead) : std::string(s0, idEnd);
    const char* sq[2] = {nullptr, nullptr};
    const char* sqe[2] = {nullptr, nullptr};
    const char* ql[2] = {nullptr, nullptr};
    const char* qle[2] = {nullptr, nullptr};
    for (unsigned m = 0; m < nMates; m++) {
        if (!ls[4 * m + 1]) { err = "EXITING because of FATAL ERROR in reads input: unexpected end of file\n"; return -STAR_EXIT_INPUT_FILES; }
        sq[m] = ls[4 * m + 1]; sqe[m] = le[4 * m + 1];
        strip(sq[m], sqe[m]);
        if (ls[4 * m + 3]) { ql[m] = ls[4 * m + 3]; qle[m] = le[4 * m + 3]; strip(ql[m], qle[m]); } else { ql[m] = qle[m] = sq[m]; }
        const uint64_t Lr = (uint64_t)(sqe[m] - sq[m]);
        if (Lr < 1) {
            err = "EXITING because of FATAL ERROR in reads input: short read sequence line: " + std::to_string(Lr) + "\nRead Name=" + readID + "\nRead Sequence=\"" + std::string(sq[m], sqe[m]) + "\"\nDEF_readNameLengthMax=50000\nDEF_readSeqLengthMax=650\n";
            return -STAR_EXIT_INPUT_FILES;
        }
        if (Lr > STAR_READ_SEQ_LENGTH_MAX) {
            err = "EXITING because of FATAL ERROR in reads input: Lread>=" + std::to_string(Lr) + "   while DEF_readSeqLengthMax=650\nRead Name=" + readID + "\nSOLUTION: increase DEF_readSeqLengthMax in IncludeDefine.h and re-compile STAR\n";
            return -STAR_EXIT_INPUT_FILES;
        }
        if ((uint64_t)(qle[m] - ql[m]) != Lr) {
            err = "EXITING because of FATAL ERROR in reads input: quality string length is not equal to sequence length\n" + readID + "\n" + std::string(sq[m], sqe[m]) + "\n" + std::string(ql[m], qle[m]) + "\nSOLUTION: fix your fastq file\n";
            return -STAR_EXIT_INPUT_FILES;
        }
    }
    if (nMates == 2 && (uint64_t)(sqe[0] - sq[0]) + (uint64_t)(sqe[1] - sq[1]) + 1 > STAR_READ_SEQ_LENGTH_MAX) {
        err = "EXITING because of FATAL ERROR in reads input: Lread of the pair = " + std::to_string((sqe[0] - sq[0]) + (sqe[1] - sq[1]) + 1) + "   while DEF_readSeqLengthMax=650\nRead Name=" + readID + "\nSOLUTION: increase DEF_readSeqLengthMax in IncludeDefine.h and re-compile STAR\n";
        return -STAR_EXIT_INPUT_FILES;
    }
    for (unsigned m = 0; m < nMates; m++) {
        if (P->clipYes) { if (c.seqOffC.empty()) c.seqOffC.push_back(0); clipMateAppend(*P, m, sq[m], (uint64_t)(sqe[m] - sq[m]), c); }
        c.seq.append(sq[m], sqe[m]);
        c.qual.append(ql[m], qle[m]);
        if (P->outQSconversionAdd != 0 && ls[4 * m + 3])   // readLoad.cpp:71-81
            for (size_t k = c.qual.size() - (size_t)(qle[m] - ql[m]); k < c.qual.size(); k++) { int qs = (int)(unsigned char)c.qual[k] + P->outQSconversionAdd; c.qual[k] = (char)(qs < 33 ? 33 : (qs > 126 ? 126 : qs)); }
        c.seqOff.push_back(c.seq.size());
    }
    {
        std::string full = readID;
        for (char sc : P->readNameSeparatorChar) {
            size_t pos = full.find(sc);
            if (pos != std::string::npos) full.resize(pos);
        }
        if (full.size() > 0) c.names.append(full, 1, std::string::npos);
    }
    c.names.push_back('\0');
    c.nameOff.push_back((uint32_t)c.names.size());
    if (P->outReadsUnmapped == "Fastx") { c.nameFullOff.push_back((uint32_t)c.namesFull.size()); c.namesFull += readID; c.namesFull.push_back('\0'); }
    c.readFilter.push_back(passFilter);
    c.iReadAll.push_back(iRead);
    c.nReads++;
    return 0;
}

// Memory-mapped chunk: (A) one thread per mate indexes the lines of the next records, (B) runThreadN threads parse contiguous
// record ranges into private pieces, (C) the pieces are concatenated in order.  Output identical to the stream parser.
long long ReadsReader::nextFast(ReadChunk& c, uint32_t maxReads, std::string& err) {
    c.clear();
    c.nMates = nMates;
    c.fastq = true;
    c.seqOff.push_back(0);
    c.nameOff.push_back(0);
    uint64_t want = maxReads;
    if (P->readMapNumber >= 0) want = std::min<uint64_t>(want, (uint64_t)P->readMapNumber > iReadAll ? (uint64_t)P->readMapNumber - iReadAll : 0);
    if (shardHi != ~0ULL) want = std::min<uint64_t>(want, shardHi > iReadAll ? shardHi - iReadAll : 0);
    if (want == 0) return 0;
    auto T0 = std::chrono::steady_clock::now();
    // (A) line index: starts[m][k], ends[m][k] for k < 4*want (shorter at the end of the file)
    std::vector<const char*>* st = lineSt_;
    std::vector<const char*>* en = lineEn_;
    for (int m = 0; m < 2; m++) { st[m].clear(); en[m].clear(); }
    size_t newOff[2] = {mapOff[0], mapOff[1]};
    long long badOff = -1;   // offset of a record start (mate 1) that is neither '@' nor end of input
    // every mate's lines are found by several threads: the range expected to hold 4*want lines (from the line length seen so far) is cut
    // into slices, each thread collects the newlines of its slice (this is also where the pages of the mapped file are faulted in), the
    // slices are concatenated in order; short of lines, the next range is scanned the same way
    const int nIdx = (int)std::max(1, std::min(16, P->stageThreads() / (int)nMates));
    auto indexMate = [&](unsigned m) {
        const char* base = map[m];
        const size_t size = mapSize[m], off0 = mapOff[m];
        const uint64_t wantLines = 4 * want;
        std::vector<const char*>& E = en[m];
        E.reserve(wantLines);
        std::vector<std::vector<const char*>>& found = idxFound_[m];
        if ((int)found.size() < nIdx) found.resize(nIdx);
        size_t scanned = off0;
        while (E.size() < wantLines && scanned < size) {
            const uint64_t need = wantLines - E.size();
            const size_t est = (size_t)((double)need * lineBytes_[m] * 1.02) + 65536;
            const size_t hi = est < size - scanned ? scanned + est : size;
            const int nT = (int)std::max<size_t>(1, std::min<size_t>((size_t)nIdx, (hi - scanned) >> 20));
            auto scan = [&](int t) {
                std::vector<const char*>& F = found[t];
                F.clear();
                const char* p = base + scanned + (hi - scanned) * (size_t)t / nT;
                const char* e = base + scanned + (hi - scanned) * (size_t)(t + 1) / nT;
                while (p < e) {
                    const char* q = (const char*)memchr(p, '\n', (size_t)(e - p));
                    if (!q) break;
                    F.push_back(q);
                    p = q + 1;
                }
            };
            if (nT == 1) scan(0);
            else {
                std::vector<std::thread> th;
                for (int t = 1; t < nT; t++) th.emplace_back(scan, t);
                scan(0);
                for (auto& t : th) t.join();
            }
            for (int t = 0; t < nT && E.size() < wantLines; t++) {
                const size_t take = std::min<size_t>(found[t].size(), wantLines - E.size());
                E.insert(E.end(), found[t].begin(), found[t].begin() + take);
            }
            scanned = hi;
        }
        if (E.size() < wantLines && (E.empty() ? off0 : (size_t)(E.back() - base) + 1) < size) E.push_back(base + size);   // last line without a newline
        std::vector<const char*>& S = st[m];
        S.resize(E.size());
        for (size_t k = 0; k < E.size(); k++) S[k] = k ? E[k - 1] + 1 : base + off0;
        if (m == 0)
            for (size_t k = 0; k < S.size(); k += 4)
                if (*S[k] != '@') {                                   // end of the records (e.g. trailing blank line) ...
                    if (*S[k] != ' ' && *S[k] != '\n') badOff = (long long)(S[k] - base);   // ... or text that is not a record: fatal in the reference
                    S.resize(k); E.resize(k);
                    break;
                }
        newOff[m] = E.empty() ? off0 : (E.back() < base + size ? (size_t)(E.back() - base) + 1 : size);
        if (!E.empty()) lineBytes_[m] = (double)(newOff[m] - off0) / (double)E.size();
    };
    if (nMates == 2) { std::thread t1(indexMate, 1u); indexMate(0); t1.join(); } else indexMate(0);
    const uint64_t nRec = (st[0].size() + 3) / 4;   // a last record with missing lines is still a record (its errors are reported)
    if (nRec == 0 && badOff >= 0) {                 // the records before the offending line are mapped first, as the reference's chunker does
        const char* b = map[0] + badOff;
        const char* q = (const char*)memchr(b, '\n', mapSize[0] - (size_t)badOff);
        err = "EXITING because of FATAL ERROR in input reads: wrong read ID line format: the read ID lines should start with @ or > \nOffending line for read # " +
              std::to_string(iReadAll + 1) + "\n" + std::string(b, q ? (size_t)(q - b) : mapSize[0] - (size_t)badOff) + "\nSOLUTION: verify and correct the input read files\n";
        return -STAR_EXIT_INPUT_FILES;
    }
    if (nRec == 0) return 0;
    if (nMates == 2 && st[1].size() > 4 * nRec) {   // mate 2 was indexed further than mate 1 has records: rewind it to the record boundary
        st[1].resize(4 * nRec); en[1].resize(4 * nRec);
        newOff[1] = (size_t)(en[1].back() - map[1]) + (en[1].back() < map[1] + mapSize[1] ? 1 : 0);
    }
    auto T1 = std::chrono::steady_clock::now();
    // (B) parallel parse
    const int nT = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)P->stageThreads(), nRec / 256 + 1));
    if ((int)parts_.size() < nT) parts_.resize(nT);
    std::vector<ReadChunk>& part = parts_;
    std::vector<std::string> perr(nT);
    std::vector<int> prc(nT, 0);
    auto work = [&](int t) {
        ReadChunk& pc = part[t];
        pc.clear(); pc.nMates = nMates; pc.seqOff.push_back(0); pc.nameOff.push_back(0);
        const uint64_t lo = nRec * t / nT, hi = nRec * (t + 1) / nT;
        pc.seq.reserve((hi - lo) * 220); pc.qual.reserve((hi - lo) * 220); pc.names.reserve((hi - lo) * 24);
        const char* ls[8]; const char* le[8];
        for (uint64_t r = lo; r < hi; r++) {
            for (unsigned m = 0; m < nMates; m++)
                for (unsigned k = 0; k < 4; k++) {
                    const uint64_t li = 4 * r + k;
                    const bool have = li < st[m].size();
                    ls[4 * m + k] = have ? st[m][li] : nullptr;
                    le[4 * m + k] = have ? en[m][li] : nullptr;
                }
            int rc = parseRecord(pc, iReadAll + r + 1, ls, le, perr[t]);
            if (rc) { prc[t] = rc; return; }
        }
    };
    if (nT == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nT; t++) th.emplace_back(work, t);
        for (auto& t : th) t.join();
    }
    for (int t = 0; t < nT; t++) if (prc[t]) { err = perr[t]; return prc[t]; }   // first error in input order
    auto T2 = std::chrono::steady_clock::now();
    // (C) concatenate
    size_t totSeq = 0, totNames = 0;
    for (int t = 0; t < nT; t++) { totSeq += part[t].seq.size(); totNames += part[t].names.size(); }
    c.seq.reserve(totSeq); c.qual.reserve(totSeq); c.names.reserve(totNames);
    c.seqOff.reserve(nRec * nMates + 1); c.nameOff.reserve(nRec + 1); c.readFilter.reserve(nRec); c.iReadAll.reserve(nRec);
    const bool plainMerge = !P->clipYes && P->outReadsUnmapped != "Fastx" && nT > 1 && getenv("STAR_B200_READER_PARALLEL_MERGE") != nullptr;   // (measured on the 128-core box: slower than the sequential appends; kept for experiments)
    if (plainMerge) {   // the common case: every parse thread copies its own piece to its final place (the copies are the reader's largest cost)
        std::vector<uint64_t> seqBase(nT + 1, 0), nameBase(nT + 1, 0), recBase(nT + 1, 0);
        for (int t = 0; t < nT; t++) { seqBase[t + 1] = seqBase[t] + part[t].seq.size(); nameBase[t + 1] = nameBase[t] + part[t].names.size(); recBase[t + 1] = recBase[t] + part[t].nReads; }
        c.seq.resize(totSeq); c.qual.resize(totSeq); c.names.resize(totNames);
        c.seqOff.resize(recBase[nT] * nMates + 1); c.nameOff.resize(recBase[nT] + 1); c.readFilter.resize(recBase[nT]); c.iReadAll.resize(recBase[nT]);
        c.seqOff[0] = 0; c.nameOff[0] = 0;
        auto merge = [&](int t) {
            const ReadChunk& pc = part[t];
            if (!pc.seq.empty()) { memcpy(&c.seq[seqBase[t]], pc.seq.data(), pc.seq.size()); memcpy(&c.qual[seqBase[t]], pc.qual.data(), pc.qual.size()); }
            if (!pc.names.empty()) memcpy(&c.names[nameBase[t]], pc.names.data(), pc.names.size());
            for (size_t k = 1; k < pc.seqOff.size(); k++) c.seqOff[recBase[t] * nMates + k] = seqBase[t] + pc.seqOff[k];
            for (size_t k = 1; k < pc.nameOff.size(); k++) c.nameOff[recBase[t] + k] = (uint32_t)(nameBase[t] + pc.nameOff[k]);
            for (size_t k = 0; k < pc.readFilter.size(); k++) c.readFilter[recBase[t] + k] = pc.readFilter[k];
            for (size_t k = 0; k < pc.iReadAll.size(); k++) c.iReadAll[recBase[t] + k] = pc.iReadAll[k];
        };
        std::vector<std::thread> th;
        for (int t = 1; t < nT; t++) th.emplace_back(merge, t);
        merge(0);
        for (auto& t : th) t.join();
        c.nReads = (uint32_t)recBase[nT];
    }
    for (int t = 0; t < nT && !plainMerge; t++) {
        ReadChunk& pc = part[t];
        const uint64_t sb = c.seq.size();
        const uint32_t nb = (uint32_t)c.names.size();
        if (P->clipYes) {
            if (c.seqOffC.empty()) c.seqOffC.push_back(0);
            const uint64_t sbC = c.seqC.size();
            c.seqC += pc.seqC;
            for (size_t k = 1; k < pc.seqOffC.size(); k++) c.seqOffC.push_back(sbC + pc.seqOffC[k]);
            c.clip5.insert(c.clip5.end(), pc.clip5.begin(), pc.clip5.end());
            c.clip3.insert(c.clip3.end(), pc.clip3.begin(), pc.clip3.end());
        }
        const uint32_t nbFull = (uint32_t)c.namesFull.size();
        c.namesFull += pc.namesFull;
        for (uint32_t o : pc.nameFullOff) c.nameFullOff.push_back(nbFull + o);
        c.seq += pc.seq; c.qual += pc.qual; c.names += pc.names;
        for (size_t k = 1; k < pc.seqOff.size(); k++) c.seqOff.push_back(sb + pc.seqOff[k]);
        for (size_t k = 1; k < pc.nameOff.size(); k++) c.nameOff.push_back(nb + pc.nameOff[k]);
        c.readFilter.insert(c.readFilter.end(), pc.readFilter.begin(), pc.readFilter.end());
        c.iReadAll.insert(c.iReadAll.end(), pc.iReadAll.begin(), pc.iReadAll.end());
        c.nReads += pc.nReads;
    }
    auto T3 = std::chrono::steady_clock::now();
    if (getenv("STAR_B200_READER_DEBUG")) fprintf(stderr, "reader: index %.1f ms, parse %.1f ms (%d threads), concat %.1f ms, %llu records\n", std::chrono::duration<double, std::milli>(T1 - T0).count(), std::chrono::duration<double, std::milli>(T2 - T1).count(), nT, std::chrono::duration<double, std::milli>(T3 - T2).count(), (unsigned long long)nRec);
    iReadAll += nRec;
    mapOff[0] = newOff[0]; mapOff[1] = newOff[1];
    return c.nReads;
}

long long ReadsReader::next(ReadChunk& c, uint32_t maxReads, std::string& err) {
    // a chunk never spans two files of a --readFilesIn list: when the current file is exhausted the next one is opened
    for (;;) {
        long long n = fast ? nextFast(c, maxReads, err) : nextStream(c, maxReads, err);
        c.fileIndex = fileIdx;
        if (n != 0) return n;
        const bool limit = (P->readMapNumber >= 0 && (long long)iReadAll >= P->readMapNumber) || iReadAll >= shardHi;
        if (limit || fileIdx + 1 >= P->readFilesNames[0].size()) return 0;
        int rc = openFile(fileIdx + 1, err);
        if (rc) return -rc;
    }
}

long long ReadsReader::nextStream(ReadChunk& c, uint32_t maxReads, std::string& err) {
    c.clear();
    c.nMates = nMates;
    c.seqOff.push_back(0);
    c.nameOff.push_back(0);
    std::string l1, seq[2], qual[2], tmp;
    while (c.nReads < maxReads) {
        if (P->readMapNumber >= 0 && (long long)iReadAll >= P->readMapNumber) break;  // processChunks.cpp:25
        if (iReadAll >= shardHi) break;                                               // end of this process' slice (multi-GPU)
        int ch = peekChar(0);
        if (ch != '@' && ch != '>') {
            if (ch == ' ' || ch == '\n' || ch < 0) break;   // end of stream (ReadAlignChunk_processChunks.cpp:192-194)
            std::string rest;                              // anything else at a record boundary is fatal there (:199-207)
            getLine(0, rest);
            err = "EXITING because of FATAL ERROR in input reads: wrong read ID line format: the read ID lines should start with @ or > \nOffending line for read # " +
                  std::to_string(iReadAll + 1) + "\n" + rest + "\nSOLUTION: verify and correct the input read files\n";
            return -STAR_EXIT_INPUT_FILES;
        }
        bool fastq = ch == '@';
        c.fastq = fastq;
        iReadAll++;
        getLine(0, l1);
        // first token = read ID; 2nd token -> Illumina filter flag (:113-127)
        size_t p0 = 0;
        while (p0 < l1.size() && !isspace((unsigned char)l1[p0])) p0++;
        std::string readID = l1.substr(0, p0);
        stripEndControl(readID);
        char passFilter = 'N';
        if (fastq && p0 < l1.size()) {
            size_t q0 = p0;
            while (q0 < l1.size() && isspace((unsigned char)l1[q0])) q0++;
            size_t q1 = q0;
            while (q1 < l1.size() && !isspace((unsigned char)l1[q1])) q1++;
            std::string field2 = l1.substr(q0, q1 - q0);
            if (field2.length() >= 3 && field2[1] == ':' && field2[2] == 'Y' && field2[3] == ':') passFilter = 'Y';
        }
        if (P->outSAMreadID == "Number") readID = std::string(fastq ? "@" : ">") + std::to_string(iReadAll);
        if (nMates == 2) getLine(1, tmp);  // the mate-2 name line is ignored (:133-135)
        for (unsigned m = 0; m < nMates; m++) {
            if (fastq) {
                if (!getLine(m, seq[m])) { err = "EXITING because of FATAL ERROR in reads input: unexpected end of file\n"; return -STAR_EXIT_INPUT_FILES; }
                stripEndControl(seq[m]);
                getLine(m, tmp);  // '+' line
                getLine(m, qual[m]);
                stripEndControl(qual[m]);
            } else {  // fasta, possibly multi-line (:158-196)
                seq[m].clear();
                for (;;) {
                    int c2 = peekChar(m);
                    if (c2 == '@' || c2 == '>' || c2 == ' ' || c2 == '\n' || c2 < 0) break;
                    getLine(m, tmp);
                    stripEndControl(tmp);
                    seq[m] += tmp;
                }
                qual[m].assign(seq[m].size(), 'A');  // readLoad.cpp:82-86
            }
            uint64_t Lr = seq[m].size();
            if (Lr < 1) {  // readLoad.cpp:38-43
                err = "EXITING because of FATAL ERROR in reads input: short read sequence line: " + std::to_string(Lr) + "\nRead Name=" + readID + "\nRead Sequence=\"" + seq[m] + "\"\nDEF_readNameLengthMax=50000\nDEF_readSeqLengthMax=650\n";
                return -STAR_EXIT_INPUT_FILES;
            }
            if (Lr > STAR_READ_SEQ_LENGTH_MAX) {  // readLoad.cpp:44-49
                err = "EXITING because of FATAL ERROR in reads input: Lread>=" + std::to_string(Lr) + "   while DEF_readSeqLengthMax=650\nRead Name=" + readID + "\nSOLUTION: increase DEF_readSeqLengthMax in IncludeDefine.h and re-compile STAR\n";
                return -STAR_EXIT_INPUT_FILES;
            }
            if (qual[m].size() != seq[m].size()) {  // readLoad.cpp:66-71
                err = "EXITING because of FATAL ERROR in reads input: quality string length is not equal to sequence length\n" + readID + "\n" + seq[m] + "\n" + qual[m] + "\nSOLUTION: fix your fastq file\n";
                return -STAR_EXIT_INPUT_FILES;
            }
        }
        if (nMates == 2 && seq[0].size() + seq[1].size() + 1 > STAR_READ_SEQ_LENGTH_MAX) {  // ReadAlign_oneRead.cpp:38-44
            err = "EXITING because of FATAL ERROR in reads input: Lread of the pair = " + std::to_string(seq[0].size() + seq[1].size() + 1) + "   while DEF_readSeqLengthMax=650\nRead Name=" + readID + "\nSOLUTION: increase DEF_readSeqLengthMax in IncludeDefine.h and re-compile STAR\n";
            return -STAR_EXIT_INPUT_FILES;
        }
        for (unsigned m = 0; m < nMates; m++) {
            if (P->clipYes) { if (c.seqOffC.empty()) c.seqOffC.push_back(0); clipMateAppend(*P, m, seq[m].data(), seq[m].size(), c); }
            c.seq += seq[m];
            c.qual += qual[m];
            if (P->outQSconversionAdd != 0 && c.fastq)
                for (size_t k = c.qual.size() - qual[m].size(); k < c.qual.size(); k++) { int qs = (int)(unsigned char)c.qual[k] + P->outQSconversionAdd; c.qual[k] = (char)(qs < 33 ? 33 : (qs > 126 ? 126 : qs)); }
            c.seqOff.push_back(c.seq.size());
        }
        // readLoad.cpp:95-98: name without the leading '@'/'>' cut at every separator character
        std::string name = readID.size() > 0 ? readID.substr(1) : std::string();
        {
            std::string full = readID;
            for (char sc : P->readNameSeparatorChar) {
                size_t pos = full.find(sc);
                if (pos != std::string::npos) full.resize(pos);
            }
            name = full.size() > 0 ? full.substr(1) : std::string();
        }
        c.names += name;
        c.names.push_back('\0');
        c.nameOff.push_back((uint32_t)c.names.size());
        if (P->outReadsUnmapped == "Fastx") { c.nameFullOff.push_back((uint32_t)c.namesFull.size()); c.namesFull += readID; c.namesFull.push_back('\0'); }
        c.readFilter.push_back(passFilter);
        c.iReadAll.push_back(iReadAll);
        c.nReads++;
    }
    return c.nReads;
}

}  // namespace starhost
