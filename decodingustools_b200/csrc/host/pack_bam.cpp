// pack_bam.cpp -- bench/test infrastructure: turn packed read columns (raw little-endian files written by numpy) into a
// coordinate-sorted single-contig BAM, BGZF-compressed on several threads.  Python's per-record writer (tests/bamio.py)
// is fine for kilobase fixtures; a chr22-size file for bench.py's e2e_bam leg needs this.
//   clb-pack-bam <dir> <contig name> <contig length> <out.bam> [threads]
// <dir> holds pos.i32 flag.u16 mapq.u8 cigar_off.u32 cigar.u32 qual_off.u64 qual.u8 name_id.u32
#include <zlib.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

template <class T> static std::vector<T> slurp(const std::string &path) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path.c_str()); exit(1); }
    fseeko(f, 0, SEEK_END); const off_t sz = ftello(f); fseeko(f, 0, SEEK_SET);
    std::vector<T> v((size_t)sz / sizeof(T));
    if (fread(v.data(), sizeof(T), v.size(), f) != v.size()) { fprintf(stderr, "short read on %s\n", path.c_str()); exit(1); }
    fclose(f);
    return v;
}
static void put32(std::vector<uint8_t> &o, uint32_t v) { o.insert(o.end(), (uint8_t *)&v, (uint8_t *)&v + 4); }

int main(int argc, char **argv) {
    if (argc < 5) { fprintf(stderr, "usage: clb-pack-bam <dir> <contig> <length> <out.bam> [threads]\n"); return 2; }
    const std::string dir = argv[1], contig = argv[2], out_path = argv[4];
    const uint32_t clen = (uint32_t)strtoul(argv[3], nullptr, 10);
    const unsigned nt = argc > 5 ? (unsigned)atoi(argv[5]) : std::max(1u, std::thread::hardware_concurrency());
    const auto pos = slurp<int32_t>(dir + "/pos.i32"); const auto flag = slurp<uint16_t>(dir + "/flag.u16"); const auto mapq = slurp<uint8_t>(dir + "/mapq.u8");
    const auto coff = slurp<uint32_t>(dir + "/cigar_off.u32"); const auto cig = slurp<uint32_t>(dir + "/cigar.u32");
    const auto qoff = slurp<uint64_t>(dir + "/qual_off.u64"); const auto qual = slurp<uint8_t>(dir + "/qual.u8"); const auto nid = slurp<uint32_t>(dir + "/name_id.u32");
    const size_t n = pos.size();
    // header
    std::vector<uint8_t> head;
    const std::string text = "@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:" + contig + "\tLN:" + std::to_string(clen) + "\n@PG\tID:bwa\tPN:bwa\n";
    head.insert(head.end(), {'B', 'A', 'M', 1}); put32(head, (uint32_t)text.size()); head.insert(head.end(), text.begin(), text.end());
    put32(head, 1); put32(head, (uint32_t)contig.size() + 1); head.insert(head.end(), contig.begin(), contig.end()); head.push_back(0); put32(head, clen);
    // records are cut into slices of ~32 MB of payload; every slice is assembled and BGZF-compressed by one thread
    std::vector<size_t> cut{0};
    {
        size_t acc = 0;
        for (size_t i = 0; i < n; i++) {
            acc += 36 + 48 + 4 * (size_t)(coff[i + 1] - coff[i]) + (size_t)(qoff[i + 1] - qoff[i]) * 3 / 2;
            if (acc >= (32u << 20)) { cut.push_back(i + 1); acc = 0; }
        }
        if (cut.back() != n) cut.push_back(n);
    }
    const size_t n_slices = cut.size() - 1;
    std::vector<std::vector<uint8_t>> comp(n_slices + 1);
    auto bgzf = [](const uint8_t *data, size_t len, std::vector<uint8_t> &out) {
        for (size_t o = 0; o < len || (len == 0 && o == 0); o += 0xff00) {
            const size_t m = std::min<size_t>(0xff00, len - o);
            uint8_t buf[0x10000 + 64];
            z_stream zs; memset(&zs, 0, sizeof zs);
            deflateInit2(&zs, 1, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
            zs.next_in = (Bytef *)(data + o); zs.avail_in = (uInt)m; zs.next_out = buf + 18; zs.avail_out = sizeof buf - 18 - 8;
            deflate(&zs, Z_FINISH);
            const size_t clen2 = zs.total_out; deflateEnd(&zs);
            const uint16_t bsize = (uint16_t)(clen2 + 25);
            const uint8_t hdr[18] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, (uint8_t)(bsize & 0xff), (uint8_t)(bsize >> 8)};
            memcpy(buf, hdr, 18);
            const uint32_t crc = (uint32_t)crc32(crc32(0, nullptr, 0), data + o, (uInt)m), isize = (uint32_t)m;
            memcpy(buf + 18 + clen2, &crc, 4); memcpy(buf + 18 + clen2 + 4, &isize, 4);
            out.insert(out.end(), buf, buf + 18 + clen2 + 8);
            if (len == 0) break;
        }
    };
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; t++)
        th.emplace_back([&, t] {
            std::vector<uint8_t> raw;
            for (size_t s = t; s < n_slices; s += nt) {
                raw.clear();
                if (s == 0) raw = head;
                char qn[96];
                for (size_t i = cut[s]; i < cut[s + 1]; i++) {
                    const int lq = snprintf(qn, sizeof qn, "A00123:7:HFLOWCELLX:1:1101:%s:%u", contig.c_str(), nid[i]) + 1;
                    const uint32_t nc = coff[i + 1] - coff[i]; const uint32_t ls = (uint32_t)(qoff[i + 1] - qoff[i]);
                    const uint32_t bs = 32 + (uint32_t)lq + 4 * nc + (ls + 1) / 2 + ls;
                    const size_t o = raw.size(); raw.resize(o + 4 + bs);
                    uint8_t *p = raw.data() + o;
                    const int32_t tid = 0, next_ref = -1, next_pos = -1, tlen = 0, ps = pos[i];
                    const uint16_t bin = 4680, ncig16 = (uint16_t)nc, fl = flag[i];
                    memcpy(p, &bs, 4); memcpy(p + 4, &tid, 4); memcpy(p + 8, &ps, 4); p[12] = (uint8_t)lq; p[13] = mapq[i];
                    memcpy(p + 14, &bin, 2); memcpy(p + 16, &ncig16, 2); memcpy(p + 18, &fl, 2); memcpy(p + 20, &ls, 4);
                    memcpy(p + 24, &next_ref, 4); memcpy(p + 28, &next_pos, 4); memcpy(p + 32, &tlen, 4);
                    memcpy(p + 36, qn, (size_t)lq);
                    memcpy(p + 36 + lq, cig.data() + coff[i], 4 * (size_t)nc);
                    memset(p + 36 + lq + 4 * nc, 0, (ls + 1) / 2);
                    memcpy(p + 36 + lq + 4 * nc + (ls + 1) / 2, qual.data() + qoff[i], ls);
                }
                bgzf(raw.data(), raw.size(), comp[s]);
            }
        });
    for (auto &x : th) x.join();
    if (n_slices == 0) bgzf(head.data(), head.size(), comp[0]);
    bgzf(nullptr, 0, comp[n_slices]);                        // EOF marker block
    FILE *f = fopen(out_path.c_str(), "wb");
    if (!f) { fprintf(stderr, "cannot create %s\n", out_path.c_str()); return 1; }
    for (auto &c : comp) fwrite(c.data(), 1, c.size(), f);
    fclose(f);
    return 0;
}
