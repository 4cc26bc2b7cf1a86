//! Golden dumps of lyon's own StrokeTessellator (lyon 1.0.x, the reference of this repo) for tests/test_lyon_golden.py.
//!
//! For every `*.path` fixture in the directory given on the command line (SVG path data, one path per file) and every
//! combination of LineJoin x LineCap x tolerance below, writes `<fixture>__<join>_<cap>_<tol>.bin`:
//!
//!   u32 magic 0x4C594F4E ("LYON"), u32 n_vertices, u32 n_indices,
//!   n_vertices x { f32 position_on_path.x, .y, f32 normal.x, .y, f32 advancement, f32 half_width, u32 src, f32 t }
//!   n_indices  x u32
//!
//! `src` = EndpointId of VertexSource::Endpoint / the `to` endpoint of VertexSource::Edge, bit 31 set on Side::Negative --
//! the 32-byte lyon_stroke_vertex of include/lyon_cuda.h.  `half_width` = line_width / 2, `t` = 1.0 for endpoints.
use lyon::extra::parser::{ParserOptions, PathParser, Source};
use lyon::path::{Path, Side};
use lyon::tessellation::*;
use std::io::Write;

#[repr(C)]
#[derive(Copy, Clone)]
struct V {
    p: [f32; 2],
    n: [f32; 2],
    adv: f32,
    hw: f32,
    src: u32,
    t: f32,
}

struct Ctor;
impl StrokeVertexConstructor<V> for Ctor {
    fn new_vertex(&mut self, v: StrokeVertex) -> V {
        let (src, t) = match v.source() {
            VertexSource::Endpoint { id } => (id.0, 1.0),
            VertexSource::Edge { to, t, .. } => (to.0, t),
        };
        let side = if v.side() == Side::Negative { 0x8000_0000u32 } else { 0 };
        V {
            p: v.position_on_path().to_array(),
            n: v.normal().to_array(),
            adv: v.advancement(),
            hw: v.line_width() * 0.5,
            src: (src & 0x7FFF_FFFF) | side,
            t,
        }
    }
}

fn main() {
    let dir = std::env::args().nth(1).expect("usage: dump_lyon_golden <dir with *.path fixtures>");
    let joins = [("miter", LineJoin::Miter), ("miterclip", LineJoin::MiterClip), ("round", LineJoin::Round), ("bevel", LineJoin::Bevel)];
    let caps = [("butt", LineCap::Butt), ("square", LineCap::Square), ("round", LineCap::Round)];
    let tols = [0.05f32, 0.1, 0.01, 0.001];
    for entry in std::fs::read_dir(&dir).unwrap() {
        let file = entry.unwrap().path();
        if file.extension().and_then(|e| e.to_str()) != Some("path") {
            continue;
        }
        let text = std::fs::read_to_string(&file).unwrap();
        let mut builder = Path::builder_with_attributes(0);
        PathParser::new()
            .parse(&ParserOptions::DEFAULT, &mut Source::new(text.chars()), &mut builder)
            .expect("fixture parses");
        let path = builder.build();
        let stem = file.file_stem().unwrap().to_str().unwrap().to_string();
        for (jn, join) in joins.iter() {
            for (cn, cap) in caps.iter() {
                for tol in tols.iter() {
                    let opts = StrokeOptions::tolerance(*tol).with_line_width(1.0).with_line_join(*join).with_line_cap(*cap);
                    let mut out: VertexBuffers<V, u32> = VertexBuffers::new();
                    StrokeTessellator::new()
                        .tessellate_path(&path, &opts, &mut BuffersBuilder::new(&mut out, Ctor))
                        .unwrap();
                    let name = format!("{}/{}__{}_{}_{}.bin", dir, stem, jn, cn, tol);
                    let mut f = std::fs::File::create(&name).unwrap();
                    f.write_all(&0x4C594F4Eu32.to_le_bytes()).unwrap();
                    f.write_all(&(out.vertices.len() as u32).to_le_bytes()).unwrap();
                    f.write_all(&(out.indices.len() as u32).to_le_bytes()).unwrap();
                    let vb = unsafe { std::slice::from_raw_parts(out.vertices.as_ptr() as *const u8, out.vertices.len() * 32) };
                    f.write_all(vb).unwrap();
                    let ib = unsafe { std::slice::from_raw_parts(out.indices.as_ptr() as *const u8, out.indices.len() * 4) };
                    f.write_all(ib).unwrap();
                    println!("{name}: {} vertices, {} indices", out.vertices.len(), out.indices.len());
                }
            }
        }
    }
}
