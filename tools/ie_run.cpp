// ie_run: load an IR with the reference Inference Engine (built by oracle/build_ref.sh), run it on a
// device ("CPU" = the MKLDNNPlugin oracle, "B200" = this repository's plugin, "MULTI:..."), feed raw
// input files and dump raw outputs + a JSON manifest.  Used by tools/make_golden.py (golden vectors)
// and by the GPU parity tests that drive the B200 plugin through the unmodified Core API
// (Core::ReadNetwork / LoadNetwork / CreateInferRequest / StartAsync / Wait).
#include <inference_engine.hpp>
#include <blob_factory.hpp>
#include <ngraph/function.hpp>
#include <ngraph/variant.hpp>

#include <chrono>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

using namespace InferenceEngine;

static std::vector<std::string> split(const std::string& s, char d) {
  std::vector<std::string> out;
  std::stringstream ss(s);
  std::string item;
  while (std::getline(ss, item, d)) out.push_back(item);
  return out;
}

int main(int argc, char** argv) {
  std::string model, device = "CPU", outdir = ".", plugins_xml;
  std::map<std::string, std::string> config, inputs, in_prec, mean_scale, mean_image;
  std::string export_path, import_path, cache_dir;
  int niter = 1;
  bool query = false, dry = false, user_blobs = false, cancel = false, callback = false, remote = false;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    auto next = [&]() { return std::string(i + 1 < argc ? argv[++i] : ""); };
    if (a == "--model") model = next();
    else if (a == "--device") device = next();
    else if (a == "--output-dir") outdir = next();
    else if (a == "--plugins") plugins_xml = next();
    else if (a == "--niter") niter = std::stoi(next());
    else if (a == "--query") query = true;     // print QueryNetwork coverage + plugin metrics as JSON, no inference
    else if (a == "--dry") dry = true;         // LoadNetwork + executable-network metrics + exec graph, no inference
    else if (a == "--export") export_path = next();   // ExecutableNetwork::Export(file) after LoadNetwork
    else if (a == "--import") import_path = next();   // Core::ImportNetwork(file, device, config) instead of ReadNetwork + LoadNetwork
    else if (a == "--cache-dir") cache_dir = next();  // Core::SetConfig({CACHE_DIR}): the Core drives Export / ImportNetwork itself
    else if (a == "--user-blobs") user_blobs = true;  // inputs and outputs in caller-owned (pageable) memory through SetBlob
    else if (a == "--remote") remote = true;          // RemoteContext + RemoteBlob inputs and outputs (device memory, filled / read through wmap / rmap)
    else if (a == "--cancel") cancel = true;          // StartAsync, Cancel, Wait: the status is written to the manifest
    else if (a == "--callback") callback = true;      // SetCompletionCallback: the number of calls is written to the manifest
    else if (a == "--config") { for (auto& kv : split(next(), ',')) { auto p = kv.find('='); if (p != std::string::npos) config[kv.substr(0, p)] = kv.substr(p + 1); } }
    else if (a == "--config-kv") { auto kv = next(); auto p = kv.find('='); if (p != std::string::npos) config[kv.substr(0, p)] = kv.substr(p + 1); }  // one key whose value has commas
    else if (a == "--input") { auto kv = next(); auto p = kv.find('='); inputs[kv.substr(0, p)] = kv.substr(p + 1); }
    else if (a == "--mean-scale") { auto kv = next(); auto p = kv.find('='); mean_scale[kv.substr(0, p)] = kv.substr(p + 1); }   // name=m0:m1:m2/s0:s1:s2 (PreProcessInfo)
    else if (a == "--mean-image") { auto kv = next(); auto p = kv.find('='); mean_image[kv.substr(0, p)] = kv.substr(p + 1); }   // name=v: MEAN_IMAGE, plane of channel c filled with v * (c + 1)
    else if (a == "--input-precision") { auto kv = next(); auto p = kv.find('='); in_prec[kv.substr(0, p)] = kv.substr(p + 1); }
    else { std::cerr << "unknown argument " << a << "\n"; return 2; }
  }
  try {
    Core core(plugins_xml);
    if (!cache_dir.empty()) core.SetConfig({{CONFIG_KEY(CACHE_DIR), cache_dir}});
    CNNNetwork net;
    if (import_path.empty()) {
      net = core.ReadNetwork(model);
      for (auto& in : net.getInputsInfo()) {
        auto it = in_prec.find(in.first);
        if (it != in_prec.end()) in.second->setPrecision(Precision::FromStr(it->second));
        auto ms = mean_scale.find(in.first);
        if (ms != mean_scale.end()) {
          auto halves = split(ms->second, '/');
          auto means = split(halves.at(0), ':'), scales = split(halves.at(1), ':');
          PreProcessInfo& pre = in.second->getPreProcess();
          pre.init(means.size());
          for (size_t c = 0; c < means.size(); ++c) { pre[c]->meanValue = std::stof(means[c]); pre[c]->stdScale = std::stof(scales.at(c)); }
          pre.setVariant(MEAN_VALUE);
        }
        auto mi = mean_image.find(in.first);
        if (mi != mean_image.end()) {
          const auto dims = in.second->getTensorDesc().getDims();
          PreProcessInfo& pre = in.second->getPreProcess();
          pre.init(dims.at(1));
          for (size_t c = 0; c < dims.at(1); ++c) {
            auto blob = make_shared_blob<float>(TensorDesc(Precision::FP32, {dims.at(2), dims.at(3)}, Layout::HW));
            blob->allocate();
            float* p = blob->buffer().as<float*>();
            for (size_t k = 0; k < blob->size(); ++k) p[k] = std::stof(mi->second) * static_cast<float>(c + 1);
            pre[c]->meanData = blob;
          }
          pre.setVariant(MEAN_IMAGE);
        }
      }
    }
    if (query) {
      QueryNetworkResult q = core.QueryNetwork(net, device, config);
      size_t total = net.getFunction()->get_ops().size();
      std::cout << "{\"supported_layers\": " << q.supportedLayersMap.size() << ", \"total_layers\": " << total;
      std::string base = device.substr(0, device.find('.'));
      auto caps = core.GetMetric(base, METRIC_KEY(OPTIMIZATION_CAPABILITIES)).as<std::vector<std::string>>();
      std::cout << ", \"capabilities\": [";
      for (size_t k = 0; k < caps.size(); ++k) std::cout << (k ? ", " : "") << "\"" << caps[k] << "\"";
      auto keys = core.GetMetric(base, METRIC_KEY(SUPPORTED_CONFIG_KEYS)).as<std::vector<std::string>>();
      std::cout << "], \"config_keys\": [";
      for (size_t k = 0; k < keys.size(); ++k) std::cout << (k ? ", " : "") << "\"" << keys[k] << "\"";
      std::cout << "], \"architecture\": \"" << core.GetMetric(base, METRIC_KEY(DEVICE_ARCHITECTURE)).as<std::string>() << "\"}" << std::endl;
      return 0;
    }
    auto t0 = std::chrono::steady_clock::now();
    RemoteContext::Ptr ctx;
    if (remote) ctx = core.GetDefaultContext(device);
    ExecutableNetwork exe = remote ? core.LoadNetwork(net, ctx, config)
                                   : (import_path.empty() ? core.LoadNetwork(net, device, config) : core.ImportNetwork(import_path, device, config));
    if (!export_path.empty()) exe.Export(export_path);
    if (dry) {
      unsigned nireq = exe.GetMetric(EXEC_NETWORK_METRIC_KEY(OPTIMAL_NUMBER_OF_INFER_REQUESTS)).as<unsigned>();
      std::string name = exe.GetMetric(EXEC_NETWORK_METRIC_KEY(NETWORK_NAME)).as<std::string>();
      CNNNetwork eg = exe.GetExecGraphInfo();
      size_t exec_nodes = 0, convs = 0;
      for (auto& n : eg.getFunction()->get_ops()) {
        auto it = n->get_rt_info().find("layerType");
        if (it == n->get_rt_info().end()) continue;
        ++exec_nodes;
        auto v = std::dynamic_pointer_cast<ngraph::VariantImpl<std::string>>(it->second);
        if (v && v->get() == "Convolution") ++convs;
      }
      std::cout << "{\"network_name\": \"" << name << "\", \"optimal_requests\": " << nireq << ", \"exec_nodes\": " << exec_nodes
                << ", \"exec_convolutions\": " << convs << ", \"input_precisions\": {";
      bool first_in = true;
      for (auto& in : exe.GetInputsInfo()) {
        std::cout << (first_in ? "" : ", ") << "\"" << in.first << "\": \"" << in.second->getPrecision().name() << "\"";
        first_in = false;
      }
      std::string prec;
      try { prec = exe.GetConfig("B200_INFERENCE_PRECISION").as<std::string>(); } catch (...) {}
      std::cout << "}, \"config_precision\": \"" << prec << "\"}" << std::endl;
      return 0;
    }
    auto t1 = std::chrono::steady_clock::now();
    InferRequest req = exe.CreateInferRequest();
    std::vector<std::vector<char>> user_mem;   // caller-owned, pageable
    if (user_blobs) {
      for (auto& in : exe.GetInputsInfo()) {
        const TensorDesc& d = in.second->getTensorDesc();
        size_t n = d.getPrecision().size();
        for (auto v : d.getDims()) n *= v;
        user_mem.emplace_back(n);
        req.SetBlob(in.first, make_blob_with_precision(d, user_mem.back().data()));
      }
      for (auto& out : exe.GetOutputsInfo()) {
        const TensorDesc& d = out.second->getTensorDesc();
        size_t n = d.getPrecision().size();
        for (auto v : d.getDims()) n *= v;
        user_mem.emplace_back(n);
        req.SetBlob(out.first, make_blob_with_precision(d, user_mem.back().data()));
      }
    }
    if (remote) {
      if (exe.GetContext()->getDeviceName() != ctx->getDeviceName()) { std::cerr << "executable network reports another context\n"; return 2; }
      for (auto& in : exe.GetInputsInfo()) req.SetBlob(in.first, ctx->CreateBlob(in.second->getTensorDesc()));
      for (auto& out : exe.GetOutputsInfo()) req.SetBlob(out.first, ctx->CreateBlob(out.second->getTensorDesc()));
    }
    int callbacks = 0;
    if (callback) req.SetCompletionCallback([&callbacks] { ++callbacks; });
    for (auto& in : exe.GetInputsInfo()) {
      Blob::Ptr b = req.GetBlob(in.first);
      auto mb = as<MemoryBlob>(b);
      auto h = mb->wmap();
      char* dst = h.as<char*>();
      auto it = inputs.find(in.first);
      if (it == inputs.end()) { std::cerr << "no file for input " << in.first << "\n"; return 2; }
      std::ifstream f(it->second, std::ios::binary);
      f.read(dst, (std::streamsize)b->byteSize());
      if ((size_t)f.gcount() != b->byteSize()) { std::cerr << "input " << in.first << ": expected " << b->byteSize() << " bytes, got " << f.gcount() << "\n"; return 2; }
    }
    std::vector<double> ms;
    int cancel_status = -100;
    for (int it = 0; it < niter; ++it) {
      auto a0 = std::chrono::steady_clock::now();
      req.StartAsync();
      if (cancel && it == 0) {
        req.Cancel();
        StatusCode st = OK;
        try { st = req.Wait(InferRequest::WaitMode::RESULT_READY); } catch (const InferCancelled&) { st = INFER_CANCELLED; }
        cancel_status = (int)st;
        req.StartAsync();     // a cancelled request must be reusable
      }
      req.Wait(InferRequest::WaitMode::RESULT_READY);
      ms.push_back(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - a0).count());
    }
    std::ofstream man(outdir + "/manifest.json");
    man << "{\"device\": \"" << device << "\", \"load_ms\": " << std::chrono::duration<double, std::milli>(t1 - t0).count() << ", \"infer_ms\": [";
    for (size_t k = 0; k < ms.size(); ++k) man << (k ? ", " : "") << ms[k];
    man << "], \"callbacks\": " << callbacks << ", \"cancel_status\": " << cancel_status << ", \"outputs\": {";
    bool first = true;
    for (auto& out : exe.GetOutputsInfo()) {
      Blob::Ptr b = req.GetBlob(out.first);
      auto mb = as<MemoryBlob>(b);
      auto h = mb->rmap();
      std::string fname = out.first;
      for (auto& ch : fname) if (ch == '/' || ch == ':') ch = '_';
      std::ofstream f(outdir + "/" + fname + ".bin", std::ios::binary);
      f.write(h.as<const char*>(), (std::streamsize)b->byteSize());
      man << (first ? "" : ", ") << "\"" << out.first << "\": {\"file\": \"" << fname << ".bin\", \"precision\": \"" << b->getTensorDesc().getPrecision().name()
          << "\", \"shape\": [";
      auto dims = b->getTensorDesc().getDims();
      for (size_t k = 0; k < dims.size(); ++k) man << (k ? ", " : "") << dims[k];
      man << "]}";
      first = false;
    }
    man << "}}\n";
  } catch (const std::exception& e) {
    std::cerr << "ie_run: " << e.what() << "\n";
    return 1;
  }
  return 0;
}
