//! Runs the reference crate on the inputs exported by `tests/golden/make_golden.py --export-inputs DIR`:
//!   DIR/<name>.pcm + DIR/<name>.desc ("channels sample_rate samples")  ->  OUT/<name>.qoa      (encode_all, src/lib.rs:760)
//!                                                                         OUT/<name>.dec.pcm  (decode_all of that, :713)
//!   DIR/<name>.qoa (foreign files, e.g. the crate's fixture)           ->  OUT/<name>.decoded.pcm
//!                                                                         OUT/<name>.reencoded.qoa (encode_all of the decode)
//! plus a streaming pass (write_header + encode_frame, :404-439) that must reproduce the one-shot bytes (:1261-1297).
//! PCM files are interleaved little-endian i16.
use qoaudio::{decode_all, encode_all, QoaDesc, QoaEncoder, QOA_FRAME_LEN};
use std::{env, fs, io::Cursor, path::Path};

fn read_pcm(p: &Path) -> Vec<i16> {
    fs::read(p).unwrap().chunks_exact(2).map(|b| i16::from_le_bytes([b[0], b[1]])).collect()
}

fn pcm_bytes(s: &[i16]) -> Vec<u8> {
    s.iter().flat_map(|v| v.to_le_bytes()).collect()
}

fn main() {
    let args: Vec<String> = env::args().collect();
    let (inp, out) = (Path::new(&args[1]), Path::new(&args[2]));
    fs::create_dir_all(out).unwrap();
    let mut names: Vec<_> = fs::read_dir(inp).unwrap().map(|e| e.unwrap().path()).collect();
    names.sort();
    for p in names {
        let stem = p.file_stem().unwrap().to_str().unwrap().to_string();
        match p.extension().and_then(|e| e.to_str()) {
            Some("pcm") => {
                let desc_txt = fs::read_to_string(p.with_extension("desc")).unwrap();
                let f: Vec<u32> = desc_txt.split_whitespace().map(|t| t.parse().unwrap()).collect();
                let desc = QoaDesc { channels: f[0] as u8, sample_rate: f[1], samples: f[2] };
                let pcm = read_pcm(&p);
                let enc = encode_all(&pcm, &desc).expect("encode_all");
                // streaming == one-shot
                let mut e = QoaEncoder::new(&desc).unwrap();
                let mut s = Vec::new();
                e.write_header(&mut s).unwrap();
                for frame in pcm.chunks(QOA_FRAME_LEN * desc.channels as usize) {
                    e.encode_frame(frame, &mut s).unwrap();
                }
                assert_eq!(s, enc, "streaming != one-shot for {stem}");
                let dec = decode_all(Cursor::new(&enc)).expect("decode_all");
                assert_eq!(dec.samples.len(), pcm.len());
                fs::write(out.join(format!("{stem}.qoa")), &enc).unwrap();
                fs::write(out.join(format!("{stem}.dec.pcm")), pcm_bytes(&dec.samples)).unwrap();
            }
            Some("qoa") => {
                let dec = decode_all(Cursor::new(fs::read(&p).unwrap())).expect("decode_all");
                fs::write(out.join(format!("{stem}.decoded.pcm")), pcm_bytes(&dec.samples)).unwrap();
                let desc = QoaDesc {
                    channels: dec.num_channels as u8,
                    sample_rate: dec.sample_rate,
                    samples: (dec.samples.len() / dec.num_channels as usize) as u32,
                };
                let enc = encode_all(&dec.samples, &desc).expect("re-encode");
                fs::write(out.join(format!("{stem}.reencoded.qoa")), &enc).unwrap();
            }
            _ => {}
        }
    }
    println!("golden-dump: done");
}
