// kuiper_tp_launch: start one process per GPU for tensor-parallel decoding with the C++ host side --
// the role torchrun plays for the Python side, without Python, MPI or a network.
//
//   kuiper_tp_launch <world> [--port P] [--all-stdout] [--] <program> [args ...]
//
// Forks <world> copies of <program>; copy r runs with KUIPER_TP_WORLD=<world>, KUIPER_TP_RANK=r,
// KUIPER_TP_PORT=P (rendezvous on 127.0.0.1; default: derived from the launcher's pid) and, unless the
// caller already set it, KUIPER_TP_DEVICE=r.  model::LLama2Model::init() reads these
// (model/tensor_parallel.h), so any program written against the kuiper:: API -- the reference's unchanged
// demo/main.cpp included -- runs sharded.  Every rank decodes the same token stream and holds the same
// logits, so only rank 0 keeps its stdout (the others go to /dev/null unless --all-stdout).  The exit
// code is the first non-zero one of the children; a failing rank takes the others down.
#include <fcntl.h>
#include <signal.h>
#include <sys/types.h>
#include <sys/wait.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {
pid_t g_kids[8];
int g_nkids = 0;
// the ranks run in process groups of their own, so a signal meant for the job has to be passed on
void forward_signal(int sig) {
  for (int k = 0; k < g_nkids; ++k) kill(-g_kids[k], sig);
}
}  // namespace

int main(int argc, char** argv) {
  if (argc < 3) {
    std::fprintf(stderr, "usage: %s <world> [--port P] [--all-stdout] [--] <program> [args ...]\n", argv[0]);
    return 2;
  }
  const int world = std::atoi(argv[1]);
  if (world != 1 && world != 2 && world != 4 && world != 8) {
    std::fprintf(stderr, "%s: world must be 1, 2, 4 or 8\n", argv[0]);
    return 2;
  }
  int port = 20000 + static_cast<int>(getpid() % 20000);
  bool all_stdout = false;
  int i = 2;
  for (; i < argc; ++i) {
    if (!std::strcmp(argv[i], "--port") && i + 1 < argc) {
      port = std::atoi(argv[++i]);
    } else if (!std::strcmp(argv[i], "--all-stdout")) {
      all_stdout = true;
    } else if (!std::strcmp(argv[i], "--")) {
      ++i;
      break;
    } else {
      break;
    }
  }
  if (i >= argc) {
    std::fprintf(stderr, "%s: no program given\n", argv[0]);
    return 2;
  }
  std::vector<pid_t> kids;
  for (int r = 0; r < world; ++r) {
    const pid_t pid = fork();
    if (pid < 0) {
      std::perror("fork");
      for (pid_t k : kids) kill(-k, SIGTERM);
      return 1;
    }
    if (pid == 0) {
      setpgid(0, 0);  // a group of its own: a failing peer takes this rank AND whatever it started down
      setenv("KUIPER_TP_WORLD", std::to_string(world).c_str(), 1);
      setenv("KUIPER_TP_RANK", std::to_string(r).c_str(), 1);
      setenv("KUIPER_TP_PORT", std::to_string(port).c_str(), 1);
      setenv("KUIPER_TP_DEVICE", std::to_string(r).c_str(), 0);
      if (r != 0 && !all_stdout) {
        const int devnull = open("/dev/null", O_WRONLY);
        if (devnull >= 0) dup2(devnull, STDOUT_FILENO);
      }
      execvp(argv[i], argv + i);
      std::perror(argv[i]);
      _exit(127);
    }
    setpgid(pid, pid);  // (both sides set it: no window in which the group does not exist yet)
    kids.push_back(pid);
    g_kids[g_nkids++] = pid;
  }
  signal(SIGINT, forward_signal);
  signal(SIGTERM, forward_signal);
  int rc = 0;
  for (size_t left = kids.size(); left > 0; --left) {
    int status = 0;
    const pid_t done = wait(&status);
    if (done < 0) break;
    const int code = WIFEXITED(status) ? WEXITSTATUS(status) : 128 + (WIFSIGNALED(status) ? WTERMSIG(status) : 0);
    if (code != 0 && rc == 0) {
      rc = code;
      for (pid_t k : kids)
        if (k != done) kill(-k, SIGTERM);  // the peers would wait a minute for the dead rank's tagged words
    }
  }
  return rc;
}
