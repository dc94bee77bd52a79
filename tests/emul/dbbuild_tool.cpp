// dbbuild_tool.cpp -- command-line wrapper of lesv_b200/host/lesv_dbbuild.hpp (the database builder the GPU
// qx2map driver uses), so that tests can compare its files with the reference's without a GPU.
#include "../../lesv_b200/host/lesv_dbbuild.hpp"
int main(int argc, char** argv)
{
    if (argc != 7) { fprintf(stderr, "usage: dbbuild_tool input db_dir title min_size max_seqs max_res\n"); return 1; }
    lesvdb::build_db(argv[1], argv[2], argv[3], atoi(argv[4]), atoi(argv[5]), (size_t)atoll(argv[6]));
    lesvdb::db_make_built(argv[1], argv[2], argv[3], atoi(argv[4]), atoi(argv[5]), (size_t)atoll(argv[6]));
    return 0;
}
