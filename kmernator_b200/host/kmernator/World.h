// World.h -- the process group of the distributed driver (FilterReads-P), one process per GPU.
// The reference gets rank/size and its collectives from Boost.MPI (mpi::communicator world, apps/FilterReads-P.cpp:263-280);
// MPI is not part of this build, so the few host-side agreements the driver needs -- the NCCL id, a common table size,
// the number of batches, the list of output files -- go through small files in a rendezvous directory, and everything on
// the data path goes through the library's communicator (kmn_comm_init).  Rank and size come from the launcher's
// environment (torchrun: RANK / WORLD_SIZE / LOCAL_RANK; mpirun: OMPI_COMM_WORLD_*; srun: SLURM_PROCID / SLURM_NTASKS).
#ifndef KMERNATOR_HOST_WORLD_H
#define KMERNATOR_HOST_WORLD_H

#include <chrono>
#include <ctime>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include <dirent.h>
#include <sys/stat.h>
#include <unistd.h>

#include "Log.h"

class World {
public:
    static World *&instance() { static World *w = NULL; return w; }      // NULL: serial run

    World() : _seq(0)
    {
        _rank = envInt("RANK", envInt("OMPI_COMM_WORLD_RANK", envInt("SLURM_PROCID", 0)));
        _size = envInt("WORLD_SIZE", envInt("OMPI_COMM_WORLD_SIZE", envInt("SLURM_NTASKS", 1)));
        _localRank = envInt("LOCAL_RANK", envInt("OMPI_COMM_WORLD_LOCAL_RANK", envInt("SLURM_LOCALID", _rank)));
        if (_size < 1 || _rank < 0 || _rank >= _size) LOG_THROW("bad rank " << _rank << " of " << _size);
        const char *d = getenv("KMN_COMM_DIR");
        std::ostringstream ss;
        if (d) ss << d;
        else {                                                             // the ranks of one launch share their parent
            const char *t = getenv("TMPDIR");
            const char *id = getenv("TORCHELASTIC_RUN_ID");
            ss << (t ? t : "/tmp") << "/kmn-comm-" << (id ? id : "run") << "-" << (long)getppid();
        }
        _dir = ss.str();
        if (_size > 1) openRendezvous();
    }
    int rank() const { return _rank; }
    int size() const { return _size; }
    int localRank() const { return _localRank; }

    // every rank contributes one string; returns all of them in rank order (a collective: same sequence on every rank)
    std::vector<std::string> allGather(const std::string &mine)
    {
        std::vector<std::string> all((size_t)_size);
        if (_size == 1) { all[0] = mine; return all; }
        const unsigned long seq = _seq++;
        const std::string fn = name(seq, _rank);
        {
            std::ofstream of((fn + ".tmp").c_str(), std::ios::binary);
            of.write(mine.data(), (std::streamsize)mine.size());
        }
        if (rename((fn + ".tmp").c_str(), fn.c_str()) != 0) LOG_THROW("could not publish " << fn);
        for (int r = 0; r < _size; ++r) {
            const std::string f = name(seq, r);
            int waited = 0;
            while (access(f.c_str(), R_OK) != 0) {
                std::this_thread::sleep_for(std::chrono::milliseconds(2));
                if (++waited > 300000) LOG_THROW("rank " << r << " did not arrive at rendezvous " << seq << " in " << _dir);
            }
            std::ifstream in(f.c_str(), std::ios::binary);
            std::ostringstream ss;
            ss << in.rdbuf();
            all[(size_t)r] = ss.str();
        }
        return all;
    }
    void barrier() { allGather(""); }
    unsigned long allMax(unsigned long v)
    {
        std::vector<std::string> a = allGather(toStr(v));
        unsigned long m = 0;
        for (size_t i = 0; i < a.size(); ++i) { unsigned long x = strtoul(a[i].c_str(), NULL, 10); if (x > m) m = x; }
        return m;
    }
    unsigned long allSum(unsigned long v)
    {
        std::vector<std::string> a = allGather(toStr(v));
        unsigned long s = 0;
        for (size_t i = 0; i < a.size(); ++i) s += strtoul(a[i].c_str(), NULL, 10);
        return s;
    }
    // rank 0's bytes to everybody (the NCCL id)
    std::string broadcast(const std::string &mine) { return allGather(_rank == 0 ? mine : std::string())[0]; }

    // end of the run: every rank says it has read everything it needs ("done.<rank>" is the last thing it touches), rank 0
    // waits for all of them and then removes the rendezvous directory
    void finalize()
    {
        if (_size == 1) return;
        barrier();
        { std::ofstream of(doneName(_rank).c_str()); of << "done"; }
        if (_rank != 0) return;
        for (int r = 0; r < _size; ++r) {
            int waited = 0;
            while (access(doneName(r).c_str(), R_OK) != 0) {
                std::this_thread::sleep_for(std::chrono::milliseconds(2));
                if (++waited > 30000) return;                              // leave the directory behind rather than hang
            }
        }
        for (unsigned long q = 0; q < _seq; ++q)
            for (int r = 0; r < _size; ++r) remove(name(q, r).c_str());
        for (int r = 0; r < _size; ++r) remove(doneName(r).c_str());
        remove((_dir + "/token").c_str());
        rmdir(_dir.c_str());
    }

private:
    // The rendezvous directory may hold the files of an earlier run that crashed (a fixed KMN_COMM_DIR, a reused parent
    // pid).  Rank 0 therefore empties it and publishes a per-launch token; every file of this launch carries the token in
    // its name, and the other ranks only accept a token that is not older than their own start (minus a launch skew of
    // ten minutes).  The directory must be ours (mode 0700), and for multi-node runs it must be on a shared file system.
    void openRendezvous()
    {
        const long t0 = (long)time(NULL);
        if (mkdir(_dir.c_str(), 0700) != 0) {
            struct stat sb;
            if (stat(_dir.c_str(), &sb) != 0 || !S_ISDIR(sb.st_mode)) LOG_THROW("cannot create the rendezvous directory " << _dir);
            if (sb.st_uid != geteuid()) LOG_THROW("rendezvous directory " << _dir << " belongs to another user");
        }
        const std::string tok = _dir + "/token";
        if (_rank == 0) {
            if (DIR *d = opendir(_dir.c_str())) {                         // leftovers of a crashed run
                while (struct dirent *e = readdir(d)) {
                    const std::string n = e->d_name;
                    if (n != "." && n != "..") remove((_dir + "/" + n).c_str());
                }
                closedir(d);
            }
            std::ostringstream ss;
            ss << t0 << "-" << (long)getpid();
            _nonce = ss.str();
            { std::ofstream of((tok + ".tmp").c_str()); of << _nonce; }
            if (rename((tok + ".tmp").c_str(), tok.c_str()) != 0) LOG_THROW("could not publish " << tok);
            return;
        }
        for (int waited = 0;; ++waited) {
            std::ifstream in(tok.c_str());
            std::string v;
            if (in.good() && (in >> v) && !v.empty() && atol(v.c_str()) >= t0 - 600) { _nonce = v; return; }
            std::this_thread::sleep_for(std::chrono::milliseconds(2));
            if (waited > 300000) LOG_THROW("rank 0 did not open the rendezvous in " << _dir);
        }
    }
    static int envInt(const char *n, int dflt) { const char *v = getenv(n); return v && *v ? atoi(v) : dflt; }
    static std::string toStr(unsigned long v) { std::ostringstream ss; ss << v; return ss.str(); }
    std::string name(unsigned long seq, int r) const { std::ostringstream ss; ss << _dir << "/" << _nonce << "." << seq << "." << r; return ss.str(); }
    std::string doneName(int r) const { std::ostringstream ss; ss << _dir << "/" << _nonce << ".done." << r; return ss.str(); }
    int _rank, _size, _localRank;
    unsigned long _seq;
    std::string _dir, _nonce;
};

#endif
