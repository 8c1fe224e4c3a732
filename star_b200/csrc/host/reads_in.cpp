// reads_in.cpp — FASTQ/FASTA input, chunked.
//
// Restates the text handling of ReadAlignChunk::processChunks (reference
// source/ReadAlignChunk_processChunks.cpp:111-196: read ID = first token of line 1, Illumina filter flag from
// the 2nd token, sequence / '+' / quality lines with one trailing control character removed) and of readLoad
// (readLoad.cpp:4-100: name cut at --readNameSeparator, length checks).  Names and qualities stay on the host;
// only the sequences go to the engine.
#include <cstring>

#include "host.h"

namespace starhost {

ReadsReader::~ReadsReader() {
    for (int m = 0; m < 2; m++)
        if (f[m]) { if (piped[m]) pclose(f[m]); else fclose(f[m]); }
}

int ReadsReader::open(const HostParams& Pin, std::string& err) {
    P = &Pin;
    nMates = Pin.readNmates;
    for (unsigned m = 0; m < nMates; m++) {
        if (Pin.readFilesCommand[0] != "-") {  // Parameters_openReadsFiles.cpp:83-101 pipes the command's stdout
            std::string cmd;
            for (auto& w : Pin.readFilesCommand) cmd += w + " ";
            cmd += "\"" + Pin.readFilesIn[m] + "\"";
            f[m] = popen(cmd.c_str(), "r");
            piped[m] = true;
        } else {
            f[m] = fopen(Pin.readFilesIn[m].c_str(), "rb");
        }
        if (!f[m]) {
            err = "EXITING because of fatal input ERROR: could not open readFilesIn=" + Pin.readFilesIn[m] + "\n";
            return STAR_EXIT_INPUT_FILES;
        }
        buf[m].resize(1 << 22);
        bpos[m] = blen[m] = 0;
    }
    if (Pin.gpuShardCount > 1) {
        // contiguous slice of the input by read index: count the records of mate 1 first (one pass over the text), then reopen
        uint64_t nRec = 0;
        {
            std::string line;
            int ch = peekChar(0);
            bool fq = ch == '@';
            uint64_t lines = 0;
            while (getLine(0, line)) { if (fq) lines++; else if (!line.empty() && line[0] == '>') nRec++; }
            if (fq) nRec = lines / 4;
        }
        for (unsigned m = 0; m < nMates; m++) { if (piped[m]) pclose(f[m]); else fclose(f[m]); f[m] = nullptr; }
        HostParams P1 = Pin;
        P1.gpuShardCount = 1;
        int rc = open(P1, err);
        P = &Pin;
        if (rc) return rc;
        shardLo = nRec * Pin.gpuShardIndex / Pin.gpuShardCount;
        shardHi = nRec * (Pin.gpuShardIndex + 1) / Pin.gpuShardCount;
        // skip the records before the slice (both mates), keeping the global read numbering
        std::string tmp;
        while (iReadAll < shardLo) {
            int ch = peekChar(0);
            if (ch != '@' && ch != '>') break;
            bool fq = ch == '@';
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

long long ReadsReader::next(ReadChunk& c, uint32_t maxReads, std::string& err) {
    c.clear();
    c.nMates = nMates;
    c.seqOff.push_back(0);
    c.nameOff.push_back(0);
    std::string l1, seq[2], qual[2], tmp;
    while (c.nReads < maxReads) {
        if (P->readMapNumber >= 0 && (long long)iReadAll >= P->readMapNumber) break;  // processChunks.cpp:25
        if (iReadAll >= shardHi) break;                                               // end of this process' slice (multi-GPU)
        int ch = peekChar(0);
        if (ch != '@' && ch != '>') break;  // end of stream (:198-200)
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
            c.seq += seq[m];
            c.qual += qual[m];
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
        c.readFilter.push_back(passFilter);
        c.iReadAll.push_back(iReadAll);
        c.nReads++;
    }
    return c.nReads;
}

}  // namespace starhost
