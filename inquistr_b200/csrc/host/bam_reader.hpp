// bam_reader.hpp -- sequential BGZF/BAM decoder on zlib (htslib is not available in this image).
// Replaces what rust-htslib does for the reference's `call` path: header @SQ names/lengths
// (call.rs:161-180) and, per record, reference_start / reference_end (bam_endpos), mapq, strand,
// CIGAR (incl. the CG:B,I long-CIGAR convention), HP and SA aux tags (call.rs:297-299,382,422-423,483).
// SEQ/QUAL are skipped, never copied.
#pragma once

#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace inqhost {

struct BamHeader {
    std::string text;
    std::vector<std::string> ref_names;
    std::vector<int64_t> ref_lens;
    int tid(const std::string &name) const;   // -1 if absent
};

// how an integer aux value was typed in the file (rust-htslib Aux variants the reference matches on)
enum class HpType : uint8_t { Absent, U8, I32, OtherInt, NotInt };

struct BamRecordView {
    int32_t tid = -1;
    int32_t pos = -1;
    int32_t end = 0;             // bam_endpos: pos + max(1, reference length); pos + 1 if unmapped
    uint8_t mapq = 0;
    uint16_t flag = 0;
    const uint32_t *cigar = nullptr;   // BAM packed words (points into the record or into the CG tag)
    uint32_t n_cigar = 0;
    HpType hp_type = HpType::Absent;
    int64_t hp_value = 0;
    bool has_sa = false;
    bool sa_is_string = false;
    std::string sa;              // SA:Z value when present
    bool is_reverse() const { return (flag & 0x10) != 0; }
};

// Streaming reader: BGZF blocks are read in batches and inflated by a small thread pool, records are
// parsed in file order.
class BamReader {
public:
    BamReader() = default;
    ~BamReader();
    BamReader(const BamReader &) = delete;
    BamReader &operator=(const BamReader &) = delete;

    // returns false and sets error() on failure
    bool open(const std::string &path, int threads);
    const BamHeader &header() const { return header_; }
    // next alignment record; false at EOF or on error (check error())
    bool next(BamRecordView &rec);
    const std::string &error() const { return err_; }
    uint64_t bytes_inflated() const { return total_out_; }

private:
    // A batch of inflated BGZF blocks. `data` starts with kSlack unused bytes so that the unconsumed
    // tail of the previous batch can be moved in front of it without copying the batch.
    struct Batch {
        std::vector<uint8_t> data;                 // capacity is recycled between batches (no re-faulting of pages)
        size_t size = 0;                           // kSlack + inflated payload bytes
        bool eof = false;
        std::string err;
    };
    std::vector<std::vector<uint8_t>> pool_;       // recycled buffers (guarded by mu_)
    static constexpr size_t kSlack = 8u << 20;
    bool next_batch();                             // make the next batch current (tail preserved)
    bool ensure_bytes(size_t n);                   // at least n unconsumed bytes are contiguous at cur_
    bool parse_header();
    void producer();                               // background thread: read + inflate batches
    bool read_batch(Batch &b);

    FILE *fp_ = nullptr;
    int threads_ = 1;
    BamHeader header_;
    std::string err_;
    Batch cur_batch_;
    size_t cur_ = 0, end_ = 0;                     // unconsumed window inside cur_batch_.data
    bool eof_ = false;
    uint64_t total_out_ = 0;
    std::vector<uint32_t> cg_;                     // aligned copy of the record's CIGAR

    std::thread producer_;
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<Batch> queue_;
    bool stop_ = false;
};

// call.rs:461-477
int64_t cigar_text_to_rlen(const std::string &cigar);
// call.rs:415-459 ; *panic set when the reference would panic (SA not a string)
bool is_accidental_2d(const BamRecordView &rec, bool *panic);

}  // namespace inqhost
