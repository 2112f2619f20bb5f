//! Dump what the real reference computes for raw RGB frames (SOURCE ONLY, see Cargo.toml).
//!
//! File names carry the parameters: `<config>_<frame>_<dictionary>_<sep>_<w>x<h>.rgb` where `<sep>` is
//! `min_corner_separation_factor` (the only non-default `DetectorConfig` field the synthetic configs use).
//! For every input `x.rgb` this writes `x.grey` (the `Detection.grey` bytes) and `x.json`
//! (`{"candidates": [[x0,y0,..,x3,y3],..], "markers": [[id, hamming_distance, code, x0,y0,..,x3,y3],..],
//!   "homography_sizes": [49,..]}`), which tools/dump_ref/compare.py diffs against the oracle's golden vectors.
use aruco3::{ARDictionary, Detector, DetectorConfig};
use std::fmt::Write as _;

fn main() {
    for path in std::env::args().skip(1) {
        let stem = std::path::Path::new(&path).file_stem().unwrap().to_str().unwrap().to_string();
        let parts: Vec<&str> = stem.split('_').collect();
        let dims: Vec<u32> = parts[parts.len() - 1].split('x').map(|v| v.parse().unwrap()).collect();
        let sep: f32 = parts[parts.len() - 2].parse().unwrap();
        let dict = parts[2..parts.len() - 2].join("_");
        let raw = std::fs::read(&path).expect("read frame");
        let img = image::RgbImage::from_raw(dims[0], dims[1], raw).expect("frame size");
        let mut config = DetectorConfig::default();
        config.min_corner_separation_factor = sep;
        let detector = Detector { config, dictionary: ARDictionary::new_from_named_dict(&dict) };
        let det = detector.detect(img.into());

        let base = path.trim_end_matches(".rgb");
        std::fs::write(format!("{base}.grey"), det.grey.as_ref().expect("grey").as_raw()).unwrap();
        let mut s = String::from("{\"candidates\":[");
        for (i, c) in det.candidates.iter().enumerate() {
            if i > 0 { s.push(','); }
            let v: Vec<String> = c.iter().flat_map(|p| [p.x.to_string(), p.y.to_string()]).collect();
            write!(s, "[{}]", v.join(",")).unwrap();
        }
        s.push_str("],\"markers\":[");
        for (i, m) in det.markers.iter().enumerate() {
            if i > 0 { s.push(','); }
            let v: Vec<String> = m.corners.iter().flat_map(|p| [p.0.to_string(), p.1.to_string()]).collect();
            write!(s, "[{},{},{},{}]", m.id, m.hamming_distance, m.code, v.join(",")).unwrap();
        }
        s.push_str("],\"homography_sizes\":[");
        let hs: Vec<String> = det.homographies.iter().map(|h| h.width().to_string()).collect();
        s.push_str(&hs.join(","));
        s.push_str("]}");
        std::fs::write(format!("{base}.json"), s).unwrap();
        println!("{stem}: {} candidates, {} markers", det.candidates.len(), det.markers.len());
    }
}
