// qx2splitseq -- cut long reads into subreads, the step in front of qx2map (reference
// src/app/split_seq/main.c:28-45, driven by scripts/x_hqx2splitseq.sh):
//     qx2splitseq seg_len ovlp_len min_last_seg_len input.fasta[.gz] > split.fasta
// A read longer than seg_len becomes pieces [from, to) named <name>_<from>_<to>; a last piece shorter than
// min_last_seg_len is merged into the one before it; consecutive pieces overlap by ovlp_len.  A read that is not
// cut keeps its name.  Names are the first token of the defline; FASTA and 4-line FASTQ input, plain or gzip;
// `input` may also be a list of file names, one per line (hbn_guess_db_format, main.c:84-100).  Host I/O only.
#include "lesv_dbbuild.hpp"

#include <string>

using namespace lesvdb;

namespace {
int g_seg = 0, g_ovlp = 0, g_min_last = 0;

void split_one(const std::string& defline, const std::string& seq)     // split_one_subject, main.c:28-45
{
    size_t n = 0;
    while (n < defline.size() && !isspace((unsigned char)defline[n])) ++n;
    const std::string name = defline.substr(0, n);
    const int len = (int)seq.size();
    int from = 0;
    while (from < len) {
        int to = from + g_seg;
        if (to > len) to = len;
        if (len - to < g_min_last) to = len;
        if (from == 0 && to == len) fprintf(stdout, ">%s\n", name.c_str());
        else fprintf(stdout, ">%s_%d_%d\n", name.c_str(), from, to);
        fwrite(seq.data() + from, 1, (size_t)(to - from), stdout);
        fputc('\n', stdout);
        from = to < len ? to - g_ovlp : len;
    }
}

bool looks_like_sequence_file(const std::string& path)
{
    LineReader in(path);
    std::string line;
    while (in.getline(line)) {
        if (line.empty()) continue;
        return line[0] == '>' || line[0] == '@';
    }
    return true;                                       // empty file: nothing to do either way
}

void split_file(const std::string& path)
{
    LineReader in(path);
    std::string line, name, seq;
    bool have = false;
    while (in.getline(line)) {
        if (line.empty()) continue;
        if (line[0] == '>') {
            if (have) split_one(name, seq);
            name = line.substr(1); seq.clear(); have = true;
        } else if (line[0] == '@' && !have) {          // FASTQ: 4-line records
            name = line.substr(1);
            if (!in.getline(seq)) break;
            std::string plus, qual;
            in.getline(plus); in.getline(qual);
            std::string clean;
            for (char c : seq) if (!isspace((unsigned char)c)) clean.push_back(c);
            split_one(name, clean);
            seq.clear();
        } else if (have) {
            for (char c : line) {
                if (c == ';') break;                   // fasta.c:213: rest of the line is a comment
                if (!isspace((unsigned char)c)) seq.push_back(c);
            }
        }
    }
    if (have) split_one(name, seq);
}
}  // namespace

int main(int argc, char** argv)
{
    if (argc != 5) {
        fprintf(stderr, "USAGE:\n%s seg_len ovlp_len min_last_seg_len input\n", argv[0]);
        return 1;
    }
    g_seg = atoi(argv[1]); g_ovlp = atoi(argv[2]); g_min_last = atoi(argv[3]);
    if (g_seg <= 0 || g_ovlp >= g_seg) die("need 0 <= ovlp_len < seg_len");
    const std::string input = argv[4];
    if (file_size(input) == 0) return 0;
    if (looks_like_sequence_file(input)) split_file(input);
    else {
        LineReader in(input);
        std::string line;
        while (in.getline(line)) {
            size_t a = 0, b = line.size();
            while (a < b && isspace((unsigned char)line[a])) ++a;
            while (b > a && isspace((unsigned char)line[b - 1])) --b;
            if (a < b) split_file(line.substr(a, b - a));
        }
    }
    return 0;
}
