// kuiper_tokenize: exercise op::SpeEncodeLayer (the tokenizer front end of model::LLama2Model) from
// the command line; the tokenizer tests compare it with the SentencePiece Python package.
//
//   kuiper_tokenize [--llama3|--qwen] <tokenizer file> <mode>   (default: SentencePiece tokenizer.model;
//                   --llama3 / --qwen: byte-level BPE tokenizer.json through Bpe/QwenEncodeLayer)
//   kuiper_tokenize <tokenizer.model> encode   < lines of text      -> one line of ids per input line
//   kuiper_tokenize <tokenizer.model> decode   < lines of ids       -> one line of hex-encoded UTF-8 text per input line
//   kuiper_tokenize <tokenizer.model> info                          -> vocab size, bos, eos
//
// encode prints what Model::encode returns for demo/main.cpp (BOS prepended, no EOS).
#include <glog/logging.h>

#include <cstdio>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "op/encode.h"

int main(int argc, char** argv) {
  std::string family = "spe";
  if (argc > 1 && std::string(argv[1]).rfind("--", 0) == 0) {
    family = argv[1] + 2;
    --argc, ++argv;
  }
  if (argc < 3) {
    std::fprintf(stderr, "usage: %s [--llama3|--qwen] <tokenizer file> encode|decode|info\n", argv[0]);
    return 2;
  }
  const std::string mode = argv[2];
  // has_bos as model.cpp:105-115 constructs them: SentencePiece and Llama-3 prepend BOS, Qwen2 does not
  std::unique_ptr<op::EncodeLayerBase> owned;
  if (family == "llama3") owned = std::make_unique<op::BpeEncodeLayer>(argv[1], true, false);
  else if (family == "qwen") owned = std::make_unique<op::QwenEncodeLayer>(argv[1], false, false);
  else owned = std::make_unique<op::SpeEncodeLayer>(argv[1], true, false);
  op::EncodeLayerBase& layer = *owned;
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
