"""myQL -- the reference's quantisation library surface (quan_func / quan_classes / graph_modify /
quan_modules), re-implemented over libsesrq.  `myQL.state` replaces the reference's ./output_pt file
protocol with an in-memory registry (files are still read / written when define.COMPAT_FILES is on)."""
