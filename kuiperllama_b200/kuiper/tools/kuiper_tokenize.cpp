// kuiper_tokenize: exercise op::SpeEncodeLayer (the tokenizer front end of model::LLama2Model) from
// the command line; the tokenizer tests compare it with the SentencePiece Python package.
//
//   kuiper_tokenize <tokenizer.model> encode   < lines of text      -> one line of ids per input line
//   kuiper_tokenize <tokenizer.model> decode   < lines of ids       -> one line of hex-encoded UTF-8 text per input line
//   kuiper_tokenize <tokenizer.model> info                          -> vocab size, bos, eos
//
// encode prints what Model::encode returns for demo/main.cpp (BOS prepended, no EOS).
#include <glog/logging.h>

#include <cstdio>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "op/encode.h"

int main(int argc, char** argv) {
  if (argc < 3) {
    std::fprintf(stderr, "usage: %s <tokenizer.model> encode|decode|info\n", argv[0]);
    return 2;
  }
  const std::string mode = argv[2];
  op::SpeEncodeLayer layer(argv[1], /*has_bos=*/true, /*has_eos=*/false);
  if (mode == "info") {
    std::printf("vocab %d eos_is_2 %d\n", layer.vocab_size(), layer.is_sentence_ending(2) ? 1 : 0);
    return 0;
  }
  std::string line;
  while (std::getline(std::cin, line)) {
    if (mode == "encode") {
      // "\\n" in the input stands for a newline inside the sentence
      std::string text;
      for (size_t i = 0; i < line.size(); ++i) {
        if (line[i] == '\\' && i + 1 < line.size() && line[i + 1] == 'n') {
          text.push_back('\n');
          ++i;
        } else {
          text.push_back(line[i]);
        }
      }
      const auto ids = layer.encode(text);
      for (size_t i = 0; i < ids.size(); ++i) std::printf("%s%d", i ? " " : "", ids[i]);
      std::printf("\n");
    } else if (mode == "decode") {
      std::istringstream is(line);
      std::vector<int32_t> ids;
      int v;
      while (is >> v) ids.push_back(v);
      const std::string text = layer.decode(ids);
      for (unsigned char c : text) std::printf("%02x", c);  // hex: the text may hold any byte
      std::printf("\n");
    } else {
      return 2;
    }
  }
  return 0;
}
